/* ref_dropin.cpp — the edit INTEGRATION.md §2 describes, as a program that compiles.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile (`make dropin`) against the UNMODIFIED
 * reference headers where they lie (-I/root/reference/raytracer_gamma: raytracer.h for
 * `struct Sphere` / `struct Light` / `struct Material` / `Vec` and the material setters) and
 * linked with raytracer-gamma_b200/librt_cuda.so, into oracle/_ref/ref_dropin.  It shows that
 * the reference's own structs, filled by the reference's own setters, go through the C-ABI
 * with nothing but pointer casts, and that the frame that comes back is the frame the
 * reference's CPU path renders: the GPU test runs the prebuilt binary and compares the PPM's
 * md5 with the reference build's (tests/golden/facts.json).
 *
 * Scene and constants are main.cpp:105-168 (values restated in a table, not pasted), the
 * quantiser is main.cpp:71-76, the PPM header main.cpp:64-66. */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <raytracer.h>      /* the reference */
#include "rt_cuda.h"        /* this repo     */

static_assert(sizeof(struct Sphere) == sizeof(rt_sphere) && sizeof(struct Light) == sizeof(rt_light) &&
              sizeof(Vec) == sizeof(rt_vec), "the reference's structs are the C-ABI's PODs");
static_assert(offsetof(struct Sphere, radius) == offsetof(rt_sphere, radius) &&
              offsetof(struct Sphere, material) == offsetof(rt_sphere, material), "Sphere layout");

#define checkRt(E, S) do { int e_ = (E); if (e_ != RT_CUDA_OK) { \
    printf("! %s: %s\n", S, rt_cuda_strerror(e_)); exit(EXIT_FAILURE); } } while (0)

int main(int argc, char** argv) {
  const unsigned kScreenWidth = 800, kScreenHeight = 600;      /* main.cpp:105-106 */
  const float zoomFactor = -4.f, aliasFactor = 3.f;            /* main.cpp:107-108 */
  const char* out = argc > 1 ? argv[1] : "testPPM.ppm";

  /* main.cpp:113-158: {matte, gloss, opacity, gloss factor, index, centre, radius} */
  struct Ball { float matte[3], gloss[3], opacity, glossFactor, index, pos[3], radius; };
  const struct Ball balls[3] = {
    {{0.4f, 0.5f, 0.7f}, {0.8f, 1.f, 0.7f}, 0.8f, 0.2f, 1.5500f, {-9.f, 0.f, -13.f}, 5.f},
    {{0.4f, 0.5f, 0.7f}, {0.8f, 1.f, 0.7f}, 0.3f, 0.95f, 1.5500f, {-4.f, 1.5f, -5.f}, 2.f},
    {{0.01f, 0.8f, 0.01f}, {0.01f, 0.8f, 0.01f}, 0.6f, 0.0f, 1.5500f, {1.f, -1.f, -7.f}, 3.f}};
  const unsigned sphNum = 3, lgtNum = 2;
  struct Sphere* hSpheres = (struct Sphere*)calloc(sphNum, sizeof(struct Sphere));
  struct Light* hLights = (struct Light*)calloc(lgtNum, sizeof(struct Light));
  for (unsigned i = 0; i < sphNum; ++i) {
    Vec matte, gloss;
    vinit(matte, balls[i].matte[0], balls[i].matte[1], balls[i].matte[2]);
    vinit(gloss, balls[i].gloss[0], balls[i].gloss[1], balls[i].gloss[2]);
    setMatOpacity(&hSpheres[i].material, balls[i].opacity);                         /* the reference's setters */
    setMatteGlossBalance(&hSpheres[i].material, balls[i].glossFactor, &matte, &gloss);
    setMatRefractivityIndex(&hSpheres[i].material, balls[i].index);
    vinit(hSpheres[i].pos, balls[i].pos[0], balls[i].pos[1], balls[i].pos[2]);
    hSpheres[i].radius = balls[i].radius;
  }
  vinit(hLights[0].pos, -45.f, 10.f, 85.f); vinit(hLights[0].col, 0.5f, 0.5f, 0.5f);   /* main.cpp:161-168 */
  vinit(hLights[1].pos, 20.f, 60.f, -5.f);  vinit(hLights[1].col, 0.5f, 0.5f, 0.5f);

  rt_cuda_ctx* ctx = NULL;
  checkRt(rt_cuda_init(0, &ctx), "Opening the CUDA device");                          /* was main.cpp:182-230 */
  checkRt(rt_cuda_upload_scene(ctx, (const rt_sphere*)hSpheres, sphNum,               /* was main.cpp:277-294 */
                               (const rt_light*)hLights, lgtNum), "Copying the scene");
  checkRt(rt_cuda_render(ctx, kScreenWidth, kScreenHeight, zoomFactor, aliasFactor,    /* was main.cpp:339-358 */
                         RTSTACK_MAXSIZE), "Enqueueing kernel");
  checkRt(rt_cuda_synchronize(ctx), "Waiting for commands to finish");                 /* was main.cpp:362     */

  Vec* pixelsIntermediate = (Vec*)calloc((size_t)kScreenHeight * kScreenWidth, sizeof(Vec));
  float maxColourValue = 1.f;
  checkRt(rt_cuda_readback(ctx, (rt_vec*)pixelsIntermediate, &maxColourValue),         /* was main.cpp:460-471 */
          "Reading the image back");
  rt_cuda_destroy(ctx);                                                                /* was main.cpp:483-489 */

  /* main.cpp:43-91 */
  FILE* f = fopen(out, "wb");
  if (!f) { printf("Can't open output file\n"); return EXIT_FAILURE; }
  fprintf(f, "P6\n%u %u\n255\n", kScreenWidth, kScreenHeight);
  for (size_t i = 0; i < (size_t)kScreenWidth * kScreenHeight; ++i) {
    const float c[3] = {pixelsIntermediate[i].x, pixelsIntermediate[i].y, pixelsIntermediate[i].z};
    unsigned char b[3];
    for (int k = 0; k < 3; ++k) b[k] = (unsigned char)(fminf(1.f, c[k]) * 255 / maxColourValue);   /* main.cpp:71-76 */
    fwrite(b, 1, 3, f);
  }
  fclose(f);
  printf("wrote %s (%ux%u, max %g)\n", out, kScreenWidth, kScreenHeight, maxColourValue);
  free(pixelsIntermediate); free(hSpheres); free(hLights);
  return 0;
}
