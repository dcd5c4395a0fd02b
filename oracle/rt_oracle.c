/* rt_oracle.c — CPU ORACLE for the raytracer-gamma per-pixel trace loop.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product path is the CUDA library
 * (raytracer-gamma_b200/csrc) and fails loudly when that is missing.
 *
 * It is a plain-C restatement of the reference's CPU copy of the algorithm
 * (paths relative to /root/reference/raytracer_gamma/):
 *   pixel loop        main.cpp:383-453 (commented-out CPU loop; same as raytrace_kernel.cl:909-972)
 *   rayTrace          raytracer.h:410-636   explicit-stack Whitted recursion
 *   calcIntersection  raytracer.h:145-194   raySphere raytracer.h:81-141
 *   shadow / matte    raytracer.h:272-309 / 313-367
 *   refraction        raytracer.h:642-815   primaryContainer :245-270  polarisedReflection :370-403
 *   reflection        raytracer.h:817-842
 *   solveQuadratic    algebra.h:22-65       max / quantise algebra.h:68-91, main.cpp:43-91
 *   stack semantics   raytraceStack.h:36-68 (push on a full stack is dropped)
 *
 * PARITY PINNING: the reference ships no tests/golden vectors for this path
 * (SURVEY.md §4).  This port is pinned against the reference's own headers
 * compiled unmodified (oracle/ref_harness.cpp -> oracle/_ref/libref_s<S>.so):
 * tests/test_oracle.py requires byte-identical float framebuffers, and
 * tests/golden/ holds vectors generated from that reference build.
 *
 * Arithmetic contract: every float operation below is a single IEEE-754
 * binary32 operation in the reference's expression order (compile with
 * -ffp-contract=off, never -ffast-math).  The two places where the reference
 * promotes to double (raytracer.h:683 and :380-393) are kept in double.
 * The one piece of undefined behaviour that changes the image — the background
 * material's uninitialised opacity (main.cpp:423-426, raytracer.h:694-697) — is
 * pinned to 0.0f, which is what -ftrivial-auto-var-init=zero gives the
 * reference build.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "rt_oracle.h"

#define ORACLE_STACK_CAP 64

typedef rt_vec V3;

static inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
/* vec.h:34-40 — component order and association are part of the contract */
static inline V3 v_add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 v_sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 v_mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 v_scale(float k, V3 b) { return v3(k * b.x, k * b.y, k * b.z); }
static inline float v_dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* vec.h:41 — 1/sqrt then three multiplies, float sqrt */
static inline V3 v_unit(V3 v) { float l = 1.f / sqrtf(v_dot(v, v)); return v_scale(l, v); }

/* raytracer.h:235-241 */
static inline int significant(V3 c) {
  const float lim = 0.001f;
  return (c.x >= lim) || (c.y >= lim) || (c.z >= lim);
}

/* The ambient medium: raytracer.h:694-697 / main.cpp:423-426 via the setters at
 * raytracer.h:62-74 with glossFactor 0 and black; opacity pinned (see header). */
static rt_material ambient_material(void) {
  rt_material m;
  float km = (float)(1.0 - (double)0.f);
  m.matteColour = v_scale(km, v3(0.f, 0.f, 0.f));
  m.glossColour = v_scale(0.f, v3(0.f, 0.f, 0.f));
  m.opacity = 0.0f;
  m.refractiveIndex = 1.00f;
  return m;
}

/* raytracer.h:81-141 */
int rt_oracle_ray_sphere(const rt_sphere* s, const rt_vec* o, const rt_vec* d, float* t) {
  const float eps = 1.0e-5f;
  int ok = 0;
  V3 disp = v_sub(*o, s->pos);
  const float a = v_dot(*d, *d);
  const float b = 2.0f * v_dot(*d, disp);
  const float c = v_dot(disp, disp) - (s->radius * s->radius);
  const float radicand = (b * b) - (4.0f * a * c);
  if (radicand >= 0.0) {
    const float root = sqrtf(radicand);
    const float denom = 2.0f * a;
    const float u0 = (-b + root) / denom;
    const float u1 = (-b - root) / denom;
    float best = 10000.f;
    if (u0 > eps && u0 < best) { best = u0; ok = 1; }
    if (u1 > eps && u1 < best) { best = u1; ok = 1; }
    *t = best;
  }
  return ok;
}

typedef struct {
  int index;       /* winning sphere, -1 if none */
  V3 point, normal;
  float sqDist;
} Hit;

/* raytracer.h:145-194.  The reference rebuilds point/normal/squaredDist on
 * every improvement; only the last improvement survives, so they are built
 * once from the winning (index, t) with the same expressions. */
static int closest_hit(const rt_sphere* sph, unsigned n, V3 o, V3 d, Hit* h,
                       rt_oracle_counters* ctr) {
  float minT = 1000.f;
  int best = -1;
  ctr->rays++;
  ctr->sphere_tests += n;
  for (unsigned i = 0; i < n; ++i) {
    float t;
    if (rt_oracle_ray_sphere(&sph[i], &o, &d, &t)) {
      if (t < minT) { minT = t; best = (int)i; }
    }
  }
  h->index = best;
  if (best >= 0) {
    V3 step = v_scale(minT, d);
    h->point = v_add(o, step);
    h->normal = v_unit(v_sub(h->point, sph[best].pos));
    V3 dist = v_scale(minT, d);
    h->sqDist = v_dot(dist, dist);
  }
  return best >= 0;
}

int rt_oracle_closest_hit(const rt_sphere* sph, unsigned n, const rt_vec* o, const rt_vec* d,
                          rt_vec* point, rt_vec* normal, float* sqDist) {
  Hit h; rt_oracle_counters c; memset(&c, 0, sizeof c);
  closest_hit(sph, n, *o, *d, &h, &c);
  if (h.index >= 0) { *point = h.point; *normal = h.normal; *sqDist = h.sqDist; }
  return h.index;
}

/* raytracer.h:245-270 — first containing sphere in index order */
int rt_oracle_primary_container(const rt_sphere* sph, unsigned n, const rt_vec* pt) {
  const float eps = 1.0e-6f;
  for (unsigned i = 0; i < n; ++i) {
    const float r = sph[i].radius + eps;
    V3 dist = v_sub(*pt, sph[i].pos);
    if (v_dot(dist, dist) <= (r * r)) return (int)i;
  }
  return -1;
}

/* raytracer.h:272-309 */
static int clear_line(const rt_sphere* sph, unsigned n, V3 a, V3 b, rt_oracle_counters* ctr) {
  V3 dir = v_sub(b, a);
  const float gap = v_dot(dir, dir);
  Hit h;
  ctr->shadow_rays++;
  if (closest_hit(sph, n, a, v_unit(dir), &h, ctr)) {
    if (h.sqDist < gap) return 0;
  }
  return 1;
}

/* raytracer.h:313-367 */
static V3 matte_light(const rt_sphere* sph, unsigned n, const rt_light* lg, unsigned nl,
                      const Hit* h, rt_oracle_counters* ctr) {
  V3 sum = v3(0.f, 0.f, 0.f);
  for (unsigned i = 0; i < nl; ++i) {
    const rt_light L = lg[i];
    if (clear_line(sph, n, h->point, L.pos, ctr)) {
      V3 dist = v_sub(L.pos, h->point);
      V3 dir = v_unit(dist);
      const float incidence = v_dot(h->normal, dir);
      if (incidence > 0.f) {
        const float d2 = v_dot(dist, dist);
        const float k = incidence / d2;
        sum = v_add(sum, v_scale(k, L.col));
      }
    }
  }
  return sum;
}

/* raytracer.h:370-403 — float products, float sum/difference, then double */
static float fresnel_term(float n1, float n2, float cosA1, float cosA2) {
  const float eps = 1.0e-6f;
  const float left = n1 * cosA1;
  const float right = n2 * cosA2;
  double num = left - right;
  double den = left + right;
  den *= den;
  if (den < eps) return 1.f;
  float refl = (float)((num * num) / den);
  if (refl > 1.f) refl = 1.f;
  return refl;
}

/* algebra.h:12-65 */
int rt_oracle_solve_quadratic(float a, float b, float c, float* roots) {
  const float tol = 0.001f;
  if (fabsf(a) < tol) {
    if (fabsf(b) < tol) return 0;
    roots[0] = -c / b;
    return 1;
  }
  const float radicand = (b * b) - (4.f * a * c);
  if (fabsf(radicand) < tol) {
    roots[0] = -b / (2.f * a);
    return 1;
  }
  const float root = sqrtf(radicand);
  const float denom = 2.0f * a;
  roots[0] = (-b + root) / denom;
  roots[1] = (-b - root) / denom;
  return 2;
}

/* raytracer.h:642-815.  `inc_*` is the incident ray with intensity already
 * scaled by the transparency (raytracer.h:496-499).  No early return on total
 * internal reflection (raytracer.h:721-730): NaNs flow through on purpose. */
static void refract(const rt_sphere* sph, unsigned n, const Hit* h, V3 inc_dir, V3 inc_int,
                    const rt_material* from, rt_material* target, float* outR,
                    V3* out_dir, V3* out_int, rt_oracle_counters* ctr) {
  float cosA1 = v_dot(inc_dir, h->normal);
  float sinA1 = 0.f;
  if (cosA1 <= -1.0) { cosA1 = -1.f; sinA1 = 0.f; }
  else if (cosA1 >= +1.f) { cosA1 = 1.f; sinA1 = 0.f; }
  else { sinA1 = (float)sqrt(1.0 - (double)(cosA1 * cosA1)); }

  const float shift = 0.01f;
  V3 probe = v_add(v_scale(shift, inc_dir), h->point);
  int inside = rt_oracle_primary_container(sph, n, &probe);
  ctr->contain_queries++;
  ctr->contain_tests += (inside >= 0) ? (uint64_t)(inside + 1) : (uint64_t)n;
  *target = (inside != -1) ? sph[inside].material : ambient_material();

  const float ratio = from->refractiveIndex / target->refractiveIndex;
  const float sinA2 = ratio * sinA1;
  /* raytracer.h:721-730: sets *outR = 1 and falls through; overwritten below */

  float roots[2];
  const int nsol = rt_oracle_solve_quadratic(1.f, (2.f * cosA1),
                                             (1.f - (1.f / (ratio * ratio))), roots);
  float bestAlign = (float)-0.1;
  V3 dir = v3(0.f, 0.f, 0.f);
  for (int i = 0; i < nsol; ++i) {
    V3 cand = v_add(inc_dir, v_scale(roots[i], h->normal));
    float align = v_dot(inc_dir, cand);
    if (align > bestAlign) { bestAlign = align; dir = cand; }
  }

  float cosA2 = sqrtf(1.f - (sinA2 * sinA2));
  if (cosA1 < 0.f) cosA2 = -cosA2;

  const float Rs = fresnel_term(from->refractiveIndex, target->refractiveIndex, cosA1, cosA2);
  const float Rp = fresnel_term(from->refractiveIndex, target->refractiveIndex, cosA2, cosA1);
  *outR = (float)((Rs + Rp) * 0.5);
  *out_int = v_scale((1.f - *outR), inc_int);
  *out_dir = dir;
}

/* raytracer.h:817-842 */
static void reflect(const Hit* h, V3 inc_dir, V3* out_o, V3* out_d) {
  const float perp = 2.f * v_dot(inc_dir, h->normal);
  V3 rd = v_unit(v_sub(inc_dir, v_scale(perp, h->normal)));
  const float shift = 0.01f;
  *out_d = rd;
  *out_o = v_add(h->point, v_scale(shift, rd));
}

typedef struct {
  V3 o, d, intensity;    /* the ray of this call                      */
  int stage;             /* 0 fresh, 1 after refraction, 2 after reflection */
  V3 colour;
  Hit hit;               /* valid from stage 0 on                     */
  rt_material hitMat;    /* copy of the struck sphere's material      */
  rt_material medium;    /* "refractiveMat": medium the ray travels in */
  float R;               /* Fresnel reflection factor from refraction */
} Frame;

typedef struct { Frame e[ORACLE_STACK_CAP]; int top, cap; } Stack;

/* raytraceStack.h:48-58 — a push onto a full stack is silently dropped */
static inline void push(Stack* s, const Frame* f, rt_oracle_counters* ctr) {
  if (!(s->top >= s->cap - 1)) {
    s->e[++s->top] = *f;
    if ((uint64_t)(s->top + 1) > ctr->max_stack) ctr->max_stack = (uint64_t)(s->top + 1);
  } else {
    ctr->dropped_pushes++;
  }
}

/* raytracer.h:410-636.  `S` plays RTSTACK_MAXSIZE.  The depth test at :458 is
 * vacuous (children always get depth 1, kMaxTraceDepth = RSIZE_MAX-1), so it
 * is not restated. */
static V3 trace(const rt_sphere* sph, unsigned n, const rt_light* lg, unsigned nl,
                V3 o, V3 d, V3 intensity, const rt_material* medium, int S,
                rt_oracle_counters* ctr) {
  V3 result = v3(0.f, 0.f, 0.f);
  Stack st; st.top = -1; st.cap = S;
  Frame cur;
  memset(&cur, 0, sizeof cur);
  cur.o = o; cur.d = d; cur.intensity = intensity;
  cur.stage = 0; cur.colour = result; cur.medium = *medium;
  push(&st, &cur, ctr);

  while (st.top >= 0) {
    cur = st.e[st.top];
    --st.top;
    ctr->pops++;
    if (cur.stage == 0) {
      if (closest_hit(sph, n, cur.o, cur.d, &cur.hit, ctr)) {
        cur.hitMat = sph[cur.hit.index].material;
        if (significant(cur.intensity)) {
          const float opacity = cur.hitMat.opacity;
          const float transparency = 1.f - opacity;
          if (opacity > 0.f) {
            V3 w = v_mul(cur.intensity, cur.hitMat.matteColour);
            w = v_scale(opacity, w);
            V3 lit = matte_light(sph, n, lg, nl, &cur.hit, ctr);
            w = v_mul(lit, w);
            cur.colour = v_add(w, cur.colour);
          }
          if (transparency > 0.f) {
            rt_material target; float R = 0.f; V3 rdir, rint;
            refract(sph, n, &cur.hit, cur.d, v_scale(transparency, cur.intensity),
                    &cur.medium, &target, &R, &rdir, &rint, ctr);
            ctr->refractions++;
            cur.R = R;
            cur.stage = 1;
            push(&st, &cur, ctr);
            Frame child; memset(&child, 0, sizeof child);
            child.o = cur.hit.point; child.d = rdir; child.intensity = rint;
            child.stage = 0; child.colour = v3(0.f, 0.f, 0.f); child.medium = target;
            push(&st, &child, ctr);
          }
          result = cur.colour;
        }
        /* hit but insignificant: `result` deliberately left stale (raytracer.h:460,538) */
      } else {
        result = v_mul(cur.intensity, cur.medium.matteColour);
      }
    } else if (cur.stage == 1) {
      cur.colour = v_add(result, cur.colour);
      V3 rc = v3(1.f, 1.f, 1.f);
      float transparency = 1.f - cur.hitMat.opacity;
      float prod = transparency * cur.R;
      rc = v_scale(prod, rc);
      rc = v_add(rc, v_scale(cur.medium.opacity, cur.hitMat.glossColour));
      rc = v_mul(cur.intensity, rc);
      if (significant(rc)) {
        V3 ro, rd;
        reflect(&cur.hit, cur.d, &ro, &rd);
        ctr->reflections++;
        cur.stage = 2;
        push(&st, &cur, ctr);
        Frame child; memset(&child, 0, sizeof child);
        child.o = ro; child.d = rd; child.intensity = rc;
        child.stage = 0; child.colour = v3(0.f, 0.f, 0.f); child.medium = cur.medium;
        push(&st, &child, ctr);
      }
      result = cur.colour;
    } else {
      cur.colour = v_add(result, cur.colour);
      result = cur.colour;
    }
  }
  return result;
}

/* main.cpp:383-453 for one pixel */
static V3 shade_pixel(const rt_sphere* sph, unsigned n, const rt_light* lg, unsigned nl,
                      unsigned W, unsigned H, float zoom, float alias, int S, unsigned gid,
                      rt_oracle_counters* ctr) {
  const float worldW = 16.f, worldH = 12.f;
  const float stepX = worldW / ((float)W);
  const float stepY = worldH / ((float)H);
  const float aspect = worldW / worldH;
  const float aliasStep = stepX / alias;
  const float samples = alias * alias;
  const float inv = 1.f / samples;
  const float px = (((float)(gid % W) - (W * 0.5f))) * stepX;
  const float py = ((H * 0.5f) - ((float)(gid / W))) * stepY;
  const rt_material bg = ambient_material();
  V3 acc = v3(0.f, 0.f, 0.f);
  for (int i = 0; i < alias; ++i) {
    for (int j = 0; j < alias; ++j) {
      float x = (px + (float)(((float)j) * aliasStep)) * aspect;
      float y = (py + (float)(((float)i) * aliasStep));
      V3 d = v_unit(v3(x, y, zoom));
      ctr->samples++;
      V3 s = trace(sph, n, lg, nl, v3(0.f, 0.f, 0.f), d, v3(1.f, 1.f, 1.f), &bg, S, ctr);
      s = v_scale(inv, s);
      acc = v_add(acc, s);
    }
  }
  return acc;
}

static void add_counters(rt_oracle_counters* a, const rt_oracle_counters* b) {
  a->rays += b->rays; a->shadow_rays += b->shadow_rays; a->sphere_tests += b->sphere_tests;
  a->contain_queries += b->contain_queries; a->contain_tests += b->contain_tests;
  a->refractions += b->refractions; a->reflections += b->reflections; a->pops += b->pops;
  a->dropped_pushes += b->dropped_pushes; a->samples += b->samples;
  if (b->max_stack > a->max_stack) a->max_stack = b->max_stack;
}

int rt_oracle_render(const rt_sphere* sph, unsigned n, const rt_light* lg, unsigned nl,
                     unsigned W, unsigned H, float zoom, float alias, int S,
                     unsigned row_begin, unsigned row_count, unsigned row_step,
                     float* out, rt_oracle_counters* ctr_out, int threads) {
  if (!out || W == 0 || H == 0 || S < 1 || S > ORACLE_STACK_CAP) return -1;
  if ((n && !sph) || (nl && !lg)) return -1;
  if (row_step == 0) row_step = 1;
  if (row_count && (uint64_t)row_begin + (uint64_t)(row_count - 1) * row_step >= H) return -1;
  rt_oracle_counters total; memset(&total, 0, sizeof total);
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  (void)threads;
#endif
#pragma omp parallel num_threads(threads)
  {
    rt_oracle_counters mine; memset(&mine, 0, sizeof mine);
#pragma omp for schedule(dynamic, 1)
    for (long k = 0; k < (long)row_count; ++k) {
      unsigned row = row_begin + (unsigned)k * row_step;
      for (unsigned x = 0; x < W; ++x) {
        V3 c = shade_pixel(sph, n, lg, nl, W, H, zoom, alias, S, row * W + x, &mine);
        float* p = out + ((size_t)k * W + x) * 3;
        p[0] = c.x; p[1] = c.y; p[2] = c.z;
      }
    }
#pragma omp critical
    add_counters(&total, &mine);
  }
  if (ctr_out) *ctr_out = total;
  return 0;
}

/* algebra.h:68-91 — NaN is skipped because `>` is false; all-black maps to 1 */
float rt_oracle_max_colour(const float* rgb, size_t npixels) {
  float m = 0.f;
  for (size_t i = 0; i < npixels * 3; ++i) if (rgb[i] > m) m = rgb[i];
  if (m == 0.f) m = 1.f;
  return m;
}

/* main.cpp:71-76: (unsigned char)(std::min(1.f,p) * 255 / max).  The cast is
 * undefined for NaN / out-of-range; g++ on x86-64 emits cvttss2si (32-bit,
 * INT_MIN on overflow/NaN) and keeps the low byte — stated explicitly here. */
static inline unsigned char quantise1(float p, float maxv) {
  float clipped = (p < 1.f) ? p : 1.f;          /* std::min(1.f, p): NaN -> 1.f */
  float v = clipped * 255 / maxv;
  int iv;
  if (v >= -2147483648.0f && v < 2147483648.0f) iv = (int)v; else iv = INT_MIN;
  return (unsigned char)(iv & 0xFF);
}

void rt_oracle_quantise(const float* rgb, size_t npixels, float maxv, unsigned char* out) {
  for (size_t i = 0; i < npixels * 3; ++i) out[i] = quantise1(rgb[i], maxv);
}

int rt_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
