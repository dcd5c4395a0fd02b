/* ref_harness.cpp — thin harness around the UNMODIFIED reference headers.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile straight from the
 * sources where they lie under /root/reference/raytracer_gamma (never copied
 * into this repo) into oracle/_ref/libref_s<S>.so.  It exists to pin the C
 * restatement in rt_oracle.c and to serve as the "reference" CPU baseline.
 *
 * The reference's CPU pixel loop is commented out in main.cpp:383-453 and
 * main.cpp itself needs OpenCL, so the loop (and only the loop) is restated
 * here around the reference's own rayTrace() (raytracer.h:410); everything it
 * calls is the reference's code.  RTSTACK_MAXSIZE has no #ifndef guard
 * (raytraceStack.h:10): for S != 6 the Makefile force-includes a sed-patched
 * temporary copy of that one header (kept outside the repo, deleted after the
 * compile), which trips the include guard before raytracer.h pulls the original.
 *
 * Flags (oracle/Makefile): -O2 -std=c++11 -ffp-contract=off
 * -ftrivial-auto-var-init=zero -DRSIZE_MAX=0x7FFFFFFF -fopenmp.  The
 * zero-init flag pins the one uninitialised read that changes the image
 * (bgMaterial.opacity, raytracer.h:573 reading main.cpp:423-426 / raytracer.h:694-697).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <raytracer.h>   /* the reference, via -I/root/reference/raytracer_gamma */

#include "rt_oracle.h"

static_assert(sizeof(Sphere) == sizeof(rt_sphere), "Sphere layout");
static_assert(sizeof(Light) == sizeof(rt_light), "Light layout");
static_assert(sizeof(Vec) == sizeof(rt_vec), "Vec layout");

extern "C" int ref_stack_size(void) { return RTSTACK_MAXSIZE; }

/* main.cpp:384-453, one pixel */
static Vec ref_pixel(Sphere* hSpheres, unsigned sphNum, Light* hLights, unsigned lgtNum,
                     const unsigned kScreenWidth, const unsigned kScreenHeight,
                     float zoomFactor, float aliasFactor, unsigned gid) {
  const float kImageWorldWidth = 16.f;
  const float kImageWorldHeight = 12.f;
  const float kRayXStep = kImageWorldWidth / ((float)kScreenWidth);
  const float kRayYStep = kImageWorldHeight / ((float)kScreenHeight);
  const float aspectRatio = kImageWorldWidth / kImageWorldHeight;
  const float kAliasFactorStepInv = kRayXStep / aliasFactor;
  const float kSamplesTot = aliasFactor * aliasFactor;
  const float kSamplesTotinv = 1.f / kSamplesTot;

  const float kPxWorldX = ((((float)(gid % kScreenWidth) - (kScreenWidth * 0.5f))) * kRayXStep);
  const float kPxWorldY = ((kScreenHeight * 0.5f) - ((float)(gid / kScreenWidth))) * kRayYStep;

  struct Ray ray; vinit(ray.origin, 0.f, 0.f, 0.f); vinit(ray.intensity, 1.f, 1.f, 1.f);
  Vec pixelCol = { 0.f, 0.f, 0.f };

  struct Material bgMaterial;
  Vec black; vinit(black, 0.f, 0.f, 0.f);
  setMatteGlossBalance(&bgMaterial, 0.f, &black, &black);
  setMatRefractivityIndex(&bgMaterial, 1.00f);

  for (int i = 0; i < aliasFactor; ++i) {
    for (int j = 0; j < aliasFactor; ++j) {
      float x = (kPxWorldX + (float)(((float)j) * kAliasFactorStepInv)) * aspectRatio;
      float y = (kPxWorldY + (float)(((float)i) * kAliasFactorStepInv));
      vinit(ray.dir, x, y, zoomFactor); vnorm(ray.dir);
      Vec currentSampleCol = rayTrace(hSpheres, sphNum, hLights, lgtNum, ray, bgMaterial, 0);
      vsmul(currentSampleCol, kSamplesTotinv, currentSampleCol);
      vadd(pixelCol, pixelCol, currentSampleCol);
    }
  }
  return pixelCol;
}

extern "C" int ref_render(const rt_sphere* spheres, unsigned sphNum,
                          const rt_light* lights, unsigned lgtNum,
                          unsigned width, unsigned height, float zoom, float aliasFactor,
                          int maxStack, unsigned row_begin, unsigned row_count, unsigned row_step,
                          float* out, rt_oracle_counters* counters, int threads) {
  if (maxStack != RTSTACK_MAXSIZE) return -2;   /* this build is fixed to one stack size */
  if (!out || width == 0 || height == 0) return -1;
  if (row_step == 0) row_step = 1;
  if (row_count && (uint64_t)row_begin + (uint64_t)(row_count - 1) * row_step >= height) return -1;
  if (counters) memset(counters, 0, sizeof *counters);   /* the reference has no counters */
  Sphere* sph = (Sphere*)spheres;   /* rayTrace takes non-const pointers; it never writes */
  Light* lgt = (Light*)lights;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  (void)threads;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (long k = 0; k < (long)row_count; ++k) {
    unsigned row = row_begin + (unsigned)k * row_step;
    for (unsigned x = 0; x < width; ++x) {
      Vec c = ref_pixel(sph, sphNum, lgt, lgtNum, width, height, zoom, aliasFactor,
                        row * width + x);
      float* p = out + ((size_t)k * width + x) * 3;
      p[0] = c.x; p[1] = c.y; p[2] = c.z;
    }
  }
  return 0;
}

extern "C" float ref_max_colour(const float* rgb, size_t npixels) {
  return maxColourValuePixelBuffer((const Vec*)rgb, npixels);   /* algebra.h:68-91 */
}

/* Unit-level pass-throughs for known-answer vectors */
extern "C" int ref_ray_sphere(const rt_sphere* s, const rt_vec* o, const rt_vec* d, float* t) {
  struct Ray r; memset(&r, 0, sizeof r);
  r.origin = *(const Vec*)o; r.dir = *(const Vec*)d;
  return raySphere((Sphere*)s, &r, t) ? 1 : 0;   /* raytracer.h:81 */
}
extern "C" int ref_primary_container(const rt_sphere* s, unsigned n, const rt_vec* pt) {
  return primaryContainer((Sphere*)s, n, (const Vec*)pt);   /* raytracer.h:245 */
}
extern "C" int ref_solve_quadratic(float a, float b, float c, float* roots) {
  return solveQuadratic(a, b, c, roots);   /* algebra.h:22 */
}
/* Materials exactly as the reference's setters build them (raytracer.h:53-74) */
extern "C" void ref_make_material(rt_material* out, const rt_vec* matte, const rt_vec* gloss,
                                  float opacity, float glossFactor, float refractiveIndex) {
  struct Material m; memset(&m, 0, sizeof m);
  setMatOpacity(&m, opacity);
  setMatteGlossBalance(&m, glossFactor, (const Vec*)matte, (const Vec*)gloss);
  setMatRefractivityIndex(&m, refractiveIndex);
  memcpy(out, &m, sizeof m);
}
