/* rt_core.cuh — the per-lane trace machine of the B200 trace loop.
 *
 * What it computes is the reference's CPU copy of the algorithm
 * (/root/reference/raytracer_gamma/raytracer.h:81-842, pixel loop
 * main.cpp:383-453); how it computes it is new:
 *
 *  - The reference recurses through an explicit stack of 168-byte snapshots
 *    (raytraceStack.h:13-68).  Here every lane runs a small state machine whose
 *    only convergence point is "run one QUERY against all N spheres".  A query
 *    is a trace ray (closest hit, raytracer.h:145), a shadow ray
 *    (raytracer.h:272) or a point-containment probe (raytracer.h:245); all three
 *    use the same fused-multiply-add sphere loop (rt_kernels.cu), so a warp
 *    stays converged through the O(N) work no matter what its lanes are doing.
 *  - A suspended call keeps 14 words (colour, the pre-computed reflected ray and
 *    its intensity, the medium) instead of a 168-byte snapshot.
 *  - Every value that can change a discrete decision or is carried into the
 *    image is evaluated with single IEEE-754 binary32 operations in the
 *    reference's expression order (the ex_* helpers below: __fmul_rn/__fadd_rn/
 *    __fdiv_rn/__fsqrt_rn never contract into FMAs), so the float framebuffer is
 *    bit-identical to the reference CPU render.  The FMA filter only decides
 *    which spheres CANNOT be hit; see Filter below for the bound.
 *
 * The file is host/device portable on purpose: tests/hostsim.cpp runs the same
 * machine lane-by-lane on the CPU to check the state machine against the
 * oracle without a GPU.  That build is test infrastructure; the product path
 * is the CUDA kernel only.
 */
#ifndef RT_CORE_CUH
#define RT_CORE_CUH

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD static inline
#endif

namespace rtg {

/* ---- exact (never contracted) IEEE binary32 / binary64 primitives ---------- */
#if defined(__CUDA_ARCH__)
RT_HD float ex_add(float a, float b) { return __fadd_rn(a, b); }
RT_HD float ex_sub(float a, float b) { return __fadd_rn(a, -b); }
RT_HD float ex_mul(float a, float b) { return __fmul_rn(a, b); }
RT_HD float ex_div(float a, float b) { return __fdiv_rn(a, b); }
RT_HD float ex_sqrt(float a) { return __fsqrt_rn(a); }
RT_HD double exd_add(double a, double b) { return __dadd_rn(a, b); }
RT_HD double exd_sub(double a, double b) { return __dadd_rn(a, -b); }
RT_HD double exd_mul(double a, double b) { return __dmul_rn(a, b); }
RT_HD double exd_div(double a, double b) { return __ddiv_rn(a, b); }
RT_HD double exd_sqrt(double a) { return __dsqrt_rn(a); }
RT_HD float fast_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
/* host build: compiled with -ffp-contract=off, so these stay separate operations */
RT_HD float ex_add(float a, float b) { return a + b; }
RT_HD float ex_sub(float a, float b) { return a - b; }
RT_HD float ex_mul(float a, float b) { return a * b; }
RT_HD float ex_div(float a, float b) { return a / b; }
RT_HD float ex_sqrt(float a) { return sqrtf(a); }
RT_HD double exd_add(double a, double b) { return a + b; }
RT_HD double exd_sub(double a, double b) { return a - b; }
RT_HD double exd_mul(double a, double b) { return a * b; }
RT_HD double exd_div(double a, double b) { return a / b; }
RT_HD double exd_sqrt(double a) { return sqrt(a); }
RT_HD float fast_fma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

struct V3 { float x, y, z; };
RT_HD V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
/* vec.h:34-41 in the reference's operand order */
RT_HD V3 vadd(V3 a, V3 b) { return mk(ex_add(a.x, b.x), ex_add(a.y, b.y), ex_add(a.z, b.z)); }
RT_HD V3 vsub(V3 a, V3 b) { return mk(ex_sub(a.x, b.x), ex_sub(a.y, b.y), ex_sub(a.z, b.z)); }
RT_HD V3 vmul(V3 a, V3 b) { return mk(ex_mul(a.x, b.x), ex_mul(a.y, b.y), ex_mul(a.z, b.z)); }
RT_HD V3 vscale(float k, V3 b) { return mk(ex_mul(k, b.x), ex_mul(k, b.y), ex_mul(k, b.z)); }
RT_HD float vdot(V3 a, V3 b) {
  return ex_add(ex_add(ex_mul(a.x, b.x), ex_mul(a.y, b.y)), ex_mul(a.z, b.z));
}
RT_HD V3 vunit(V3 v) { float l = ex_div(1.f, ex_sqrt(vdot(v, v))); return vscale(l, v); }
/* raytracer.h:235-241 (NaN is not significant) */
RT_HD bool significant(V3 c) { return (c.x >= 0.001f) || (c.y >= 0.001f) || (c.z >= 0.001f); }

/* ---- scene as the kernels see it (SoA, built once by the shim) -------------- */
struct float4_ { float x, y, z, w; };   /* plain 16-byte record, same layout as CUDA float4 */

struct SceneView {
  const float4_* filt;   /* [nPad] {cx, cy, cz, w}: filter record, see Filter         */
  const float4_* geo;    /* [n]    {cx, cy, cz, radius}: exact geometry                */
  const float4_* matA;   /* [n+1]  {matte.rgb, opacity}; entry n is the ambient medium */
  const float4_* matB;   /* [n+1]  {gloss.rgb, refractiveIndex}                        */
  const float4_* lpos;   /* [nl]   {pos.xyz, 0}                                        */
  const float4_* lcol;   /* [nl]   {col.rgb, 0}                                        */
  uint32_t n, nPad, nl;
};

/* ---- Filter ---------------------------------------------------------------
 * For a query with origin o and direction d (|d|^2 = A) and sphere (c, r), the
 * reference's float radicand (raytracer.h:95-105) approximates
 *     D = 4 [ (d.delta)^2 - A (|delta|^2 - r^2) ],   delta = o - c.
 * The filter evaluates, with FMAs and d pre-scaled to unit length,
 *     lhs = (d'.(o-c))^2 - q            q  = |o|^2 (1-kappa)
 *     ch  = w - 2 o.c                    w  = |c|^2 - r^2 - kappa(|c|^2+r^2) - 2.5e-6 r - 1e-11
 * and declares a CERTAIN MISS iff lhs < ch, i.e. iff
 *     (d'.delta)^2 < |delta|^2 - r^2 - kappa(|o|^2+|c|^2+r^2) - (container slack).
 * kappa = 2^-17 = 128 u covers (first-order, u = 2^-24): 17u A(|delta|^2+r^2)
 * rounding of the reference's own non-fused radicand, 4u for its rounded delta,
 * 24u(|o|^2+|c|^2) for the expanded-form d'.(o-c) squared, 10u(|o|^2+|c|^2+r^2)
 * for ch, 8u|delta|^2 for the unit scaling — about 92u(|o|^2+|c|^2)+31u r^2 in
 * total, using |delta|^2 <= 2(|o|^2+|c|^2).  Anything not a certain miss goes
 * through the reference's exact expressions, so a loose bound costs time, never
 * correctness; a NaN anywhere compares false and is therefore kept.
 * With d' = 0 the same records answer the containment probe of
 * raytracer.h:245-270: lhs < ch  <=>  |p-c|^2 > r^2 + slack, and the slack
 * exceeds (r+1e-6)^2 - r^2.
 */
#define RT_KAPPA 7.62939453125e-06f   /* 2^-17 */

struct Query {
  float ndx, ndy, ndz, od;   /* -d', d'.o  (d' = d/|d|)                       */
  float px, py, pz, q;       /* -2o, |o|^2(1-kappa); q=+inf never flags       */
  float dx, dy, dz;          /* exact direction for the ray kinds             */
};

enum { K_NULL = 0, K_TRACE = 1, K_SHADOW = 2, K_CONTAIN = 3 };

RT_HD void query_null(Query& qy) {
  qy.ndx = qy.ndy = qy.ndz = 0.f; qy.od = 0.f;
  qy.px = qy.py = qy.pz = 0.f; qy.q = INFINITY;
  qy.dx = qy.dy = qy.dz = 0.f;
}

/* Ray query.  Returns false when d == 0: the reference then computes
 * a = b = 0, radicand = 0, u = 0/0 = NaN for every sphere (raytracer.h:98-116),
 * i.e. a guaranteed miss, so the sphere loop is skipped for that lane. */
RT_HD bool query_ray(Query& qy, V3 o, V3 d) {
  const float A = vdot(d, d);
  qy.dx = d.x; qy.dy = d.y; qy.dz = d.z;
  qy.px = -2.f * o.x; qy.py = -2.f * o.y; qy.pz = -2.f * o.z;
  if (!(A > 0.f) && !(A != A)) {   /* A == 0 (not NaN) */
    qy.ndx = qy.ndy = qy.ndz = 0.f; qy.od = 0.f; qy.q = INFINITY;
    return false;
  }
  const float s = ex_div(1.f, ex_sqrt(A));
  const float ux = ex_mul(d.x, s), uy = ex_mul(d.y, s), uz = ex_mul(d.z, s);
  qy.ndx = -ux; qy.ndy = -uy; qy.ndz = -uz;
  qy.od = fast_fma(uz, o.z, fast_fma(uy, o.y, ex_mul(ux, o.x)));
  const float oo = fast_fma(o.z, o.z, fast_fma(o.y, o.y, ex_mul(o.x, o.x)));
  qy.q = ex_mul(oo, 1.f - RT_KAPPA);
  return true;
}

RT_HD void query_point(Query& qy, V3 p) {
  qy.ndx = qy.ndy = qy.ndz = 0.f; qy.od = 0.f;
  qy.dx = qy.dy = qy.dz = 0.f;
  qy.px = -2.f * p.x; qy.py = -2.f * p.y; qy.pz = -2.f * p.z;
  const float oo = fast_fma(p.z, p.z, fast_fma(p.y, p.y, ex_mul(p.x, p.x)));
  qy.q = ex_mul(oo, 1.f - RT_KAPPA);
}

RT_HD V3 query_origin(const Query& qy) { return mk(-0.5f * qy.px, -0.5f * qy.py, -0.5f * qy.pz); }

/* One filter test.  Returns lhs - ch: NEGATIVE (sign bit set) means certain miss; zero
 * or positive means the sphere may be hit / may contain the point.  lhs and ch are finite
 * or -inf/+inf by construction (null queries, padding), and the rounded difference of two
 * such values carries the sign of the exact difference, so "sign bit set" <=> lhs < ch.
 * Queries with NaN/inf geometry never reach the filter (query_filterable). */
RT_HD float filter_diff(const Query& qy, float4_ s) {
  float bq = fast_fma(qy.ndx, s.x, qy.od);
  bq = fast_fma(qy.ndy, s.y, bq);
  bq = fast_fma(qy.ndz, s.z, bq);
  float ch = fast_fma(qy.px, s.x, s.w);
  ch = fast_fma(qy.py, s.y, ch);
  ch = fast_fma(qy.pz, s.z, ch);
  const float lhs = fast_fma(bq, bq, -qy.q);
  return ex_sub(lhs, ch);
}
RT_HD bool filter_pass(const Query& qy, float4_ s) { return !(filter_diff(qy, s) < 0.f); }

/* The sign-bit form of the filter is only conservative for non-NaN inputs.  A query whose
 * numbers are not all finite (q = +inf marks a null query and is fine) skips the filter and
 * is tested exactly against every sphere. */
RT_HD bool finite_f(float x) { return fabsf(x) <= 3.402823466e+38f; }
RT_HD bool query_filterable(const Query& qy) {
  return finite_f(qy.ndx) && finite_f(qy.ndy) && finite_f(qy.ndz) && finite_f(qy.od) &&
         finite_f(qy.px) && finite_f(qy.py) && finite_f(qy.pz) && !(qy.q != qy.q) && qy.q > -INFINITY;
}

/* ---- exact per-candidate tests --------------------------------------------- */
/* raytracer.h:81-141.  The reference forms both roots u0 = (-b+root)/denom and
 * u1 = (-b-root)/denom and keeps the smallest one in (1e-5, 10000).  Rounding is
 * monotonic and denom = 2 d.d > 0, so u1 <= u0 always: when u1 > 1e-5 it is the
 * answer (or, if u1 >= 10000, so is u0 and there is none) and u0 need not be
 * divided out; u0 is only formed when u1 fails the lower bound (or is NaN).
 * Same decisions, same t, one IEEE division less on the common path. */
RT_HD bool ray_sphere_exact(float4_ g, V3 o, V3 d, float& t) {
  const V3 disp = vsub(o, mk(g.x, g.y, g.z));
  const float a = vdot(d, d);
  const float b = ex_mul(2.0f, vdot(d, disp));
  const float c = ex_sub(vdot(disp, disp), ex_mul(g.w, g.w));
  const float radicand = ex_sub(ex_mul(b, b), ex_mul(ex_mul(4.0f, a), c));
  if (!(radicand >= 0.0f)) return false;
  const float root = ex_sqrt(radicand);
  const float denom = ex_mul(2.0f, a);
  const float u1 = ex_div(ex_sub(-b, root), denom);
  if (u1 > 1.0e-5f) {
    if (u1 < 10000.f) { t = u1; return true; }
    return false;
  }
  const float u0 = ex_div(ex_add(-b, root), denom);
  if (u0 > 1.0e-5f && u0 < 10000.f) { t = u0; return true; }
  return false;
}

/* raytracer.h:259-266 */
RT_HD bool contains_exact(float4_ g, V3 p) {
  const float r = ex_add(g.w, 1.0e-6f);
  const V3 dist = vsub(p, mk(g.x, g.y, g.z));
  return vdot(dist, dist) <= ex_mul(r, r);
}

/* ---- suspended calls -------------------------------------------------------- */
#define RT_MAX_STACK 16
struct Frame {            /* 16 words: one 64-byte local-memory record */
  V3 colour;              /* cur.colour at suspension                                  */
  float stage;            /* 1.f: waiting for the refracted child, 2.f: reflected child */
  V3 reflCol; float medium;   /* reflection intensity (raytracer.h:563-578), medium id  */
  V3 reflO; float pad0;   /* pre-computed reflected ray (raytracer.h:817-842)           */
  V3 reflD; float pad1;
};

struct Counters {         /* per-lane tallies, reduced per block at the end */
  uint32_t rays, shadow, containQ, containT, exactTests, samples;
  uint32_t nullRays;      /* rays with d == 0: answered without running the sphere loop */
};

/* ---- the lane ---------------------------------------------------------------- */
struct Lane {
  Query qy;
  int kind;
  /* result of the sphere loop */
  float minT; int hitIdx;
  /* the call being evaluated ("currSnapshot") */
  V3 rayD, rayI, colour;
  int medium;             /* sphere index of the medium the ray travels in, n = ambient */
  int obj;                /* struck sphere */
  V3 P, Nrm;
  /* matte accumulation */
  int light; V3 lit; float gap;
  /* "colourSum" of raytracer.h:425 */
  V3 result;
  int top;                /* frames on the stack = depth of the current call */
  /* pixel / sample bookkeeping */
  int si, sj; V3 acc; float pxw, pyw;
  uint32_t pixel;         /* destination index, 0xFFFFFFFF = lane has no pixel */
  Counters ctr;
};

struct Camera {           /* main.cpp:384-402, evaluated once on the host in float */
  uint32_t W, H;
  float zoom, alias;
  float stepX, stepY, aspect, aliasStep, inv;
  int nIter;              /* iterations of `for (int i = 0; i < alias; ++i)` */
  int S;                  /* RTSTACK_MAXSIZE */
  int ambient;            /* material index of the ambient medium (= sphere count) */
};

RT_HD Camera make_camera(uint32_t W, uint32_t H, float zoom, float alias, int S, int ambient) {
  Camera c;
  c.W = W; c.H = H; c.zoom = zoom; c.alias = alias; c.S = S; c.ambient = ambient;
  c.stepX = ex_div(16.f, (float)W);
  c.stepY = ex_div(12.f, (float)H);
  c.aspect = ex_div(16.f, 12.f);
  c.aliasStep = ex_div(c.stepX, alias);
  c.inv = ex_div(1.f, ex_mul(alias, alias));
  int n = 0;
  while (n < (1 << 20) && (float)n < alias) ++n;
  c.nIter = n;
  return c;
}

/* Start the sample (si, sj) of the lane's pixel: main.cpp:432-441 */
RT_HD void start_sample(Lane& L, const Camera& cam) {
  const float x = ex_mul(ex_add(L.pxw, ex_mul((float)L.sj, cam.aliasStep)), cam.aspect);
  const float y = ex_add(L.pyw, ex_mul((float)L.si, cam.aliasStep));
  const V3 d = vunit(mk(x, y, cam.zoom));
  L.rayD = d; L.rayI = mk(1.f, 1.f, 1.f); L.colour = mk(0.f, 0.f, 0.f);
  L.result = mk(0.f, 0.f, 0.f);
  L.medium = cam.ambient;                  /* every sample starts in the ambient medium, main.cpp:439 */
  L.top = -1;
  L.ctr.samples++;
  L.kind = K_TRACE;
  L.minT = 1000.f; L.hitIdx = -1;
  query_ray(L.qy, mk(0.f, 0.f, 0.f), d);   /* d is unit: never the null case */
}

/* Give the lane pixel (gx, gy) of the full frame, stored at dst index `pixel`. */
RT_HD bool start_pixel(Lane& L, const Camera& cam, uint32_t gx, uint32_t gy, uint32_t pixel) {
  L.pixel = pixel;
  L.pxw = ex_mul(ex_sub((float)gx, ex_mul((float)cam.W, 0.5f)), cam.stepX);
  L.pyw = ex_mul(ex_sub(ex_mul((float)cam.H, 0.5f), (float)gy), cam.stepY);
  L.acc = mk(0.f, 0.f, 0.f);
  L.si = 0; L.sj = 0;
  if (cam.nIter <= 0) { L.kind = K_NULL; query_null(L.qy); return false; }
  start_sample(L, cam);
  return true;
}

RT_HD void set_ray_query(Lane& L, int kind, V3 o, V3 d) {
  L.kind = kind;
  L.minT = 1000.f; L.hitIdx = -1;
  if (!query_ray(L.qy, o, d)) L.ctr.nullRays++;
}

/* raytracer.h:272-286: shadow ray towards light `L.light` */
RT_HD void setup_shadow(Lane& L, const SceneView& sc) {
  const float4_ lp = sc.lpos[L.light];
  const V3 dir = vsub(mk(lp.x, lp.y, lp.z), L.P);
  L.gap = vdot(dir, dir);
  set_ray_query(L, K_SHADOW, L.P, vunit(dir));
}

/* Sample finished: main.cpp:443-447.  Returns true when the pixel is complete. */
RT_HD bool finish_sample(Lane& L, const Camera& cam) {
  const V3 s = vscale(cam.inv, L.result);
  L.acc = vadd(L.acc, s);
  if (++L.sj >= cam.nIter) { L.sj = 0; ++L.si; }
  if (L.si >= cam.nIter) return true;
  start_sample(L, cam);
  return false;
}

/* Pop suspended calls until one launches a child ray or the stack is empty
 * (raytracer.h:552-628).  Returns true when the sample is finished. */
RT_HD bool unwind(Lane& L, Frame* stack, const Camera& cam) {
  while (L.top >= 0) {
    Frame& f = stack[L.top];
    --L.top;
    L.colour = vadd(L.result, f.colour);
    if (f.stage == 1.f) {
      if (significant(f.reflCol)) {
        /* re-push as stage 2 (always fits: the slot was just vacated) */
        ++L.top;
        f.colour = L.colour; f.stage = 2.f;
        L.result = L.colour;
        if (L.top < cam.S - 1) {
          /* reflected child, raytracer.h:602-611 */
          L.rayD = f.reflD; L.rayI = f.reflCol; L.medium = (int)f.medium;
          L.colour = mk(0.f, 0.f, 0.f);
          set_ray_query(L, K_TRACE, f.reflO, f.reflD);
          return false;
        }
        /* child push dropped (raytraceStack.h:52): fall through to pop stage 2 */
        continue;
      }
      L.result = L.colour;
    } else {
      L.result = L.colour;
    }
  }
  return true;
}

/* raytracer.h:370-403 */
RT_HD float fresnel_term(float n1, float n2, float cosA1, float cosA2) {
  const float left = ex_mul(n1, cosA1);
  const float right = ex_mul(n2, cosA2);
  const double num = (double)ex_sub(left, right);
  double den = (double)ex_add(left, right);
  den = exd_mul(den, den);
  if (den < (double)1.0e-6f) return 1.f;
  float refl = (float)exd_div(exd_mul(num, num), den);
  if (refl > 1.f) refl = 1.f;
  return refl;
}

/* What a handler leaves for the common tail of advance(): each tail (shadow set-up,
 * matte/transparency step, unwinding) exists once, so lanes of different kinds
 * re-converge on it and the code stays small. */
enum { ACT_QUERY = 0, ACT_UNWIND = 1, ACT_SHADOW = 2, ACT_MATTE = 3, ACT_NOMATTE = 4 };

/* After the containment probe: refraction (raytracer.h:642-815), suspend the
 * call (raytracer.h:516-522), launch the refracted child (raytracer.h:524-533). */
RT_HD int after_contain(Lane& L, const SceneView& sc, Frame* stack, const Camera& cam) {
  const int target = (L.hitIdx >= 0) ? L.hitIdx : (int)sc.n;
  const float4_ objA = sc.matA[L.obj], objB = sc.matB[L.obj];
  const float4_ medA = sc.matA[L.medium], medB = sc.matB[L.medium];
  const float4_ tgtB = sc.matB[target];
  const float opacity = objA.w;
  const float transparency = ex_sub(1.f, opacity);
  const V3 incI = vscale(transparency, L.rayI);          /* raytracer.h:498 */

  float cosA1 = vdot(L.rayD, L.Nrm);
  float sinA1 = 0.f;
  if (cosA1 <= -1.0f) { cosA1 = -1.f; }
  else if (cosA1 >= 1.f) { cosA1 = 1.f; }
  else { sinA1 = (float)exd_sqrt(exd_sub(1.0, (double)ex_mul(cosA1, cosA1))); }   /* :683, double */

  const float n1 = medB.w, n2 = tgtB.w;
  const float ratio = ex_div(n1, n2);
  const float sinA2 = ex_mul(ratio, sinA1);

  /* solveQuadratic(1, 2cosA1, 1 - 1/ratio^2), algebra.h:22-65 with a = 1 */
  const float qb = ex_mul(2.f, cosA1);
  const float qc = ex_sub(1.f, ex_div(1.f, ex_mul(ratio, ratio)));
  const float radicand = ex_sub(ex_mul(qb, qb), ex_mul(ex_mul(4.f, 1.f), qc));
  float r0, r1; int nsol;
  if (fabsf(radicand) < 0.001f) { r0 = ex_div(-qb, ex_mul(2.f, 1.f)); r1 = 0.f; nsol = 1; }
  else {
    const float root = ex_sqrt(radicand);
    const float denom = ex_mul(2.0f, 1.f);
    r0 = ex_div(ex_add(-qb, root), denom);
    r1 = ex_div(ex_sub(-qb, root), denom);
    nsol = 2;
  }
  float bestAlign = (float)-0.1;
  V3 rdir = mk(0.f, 0.f, 0.f);
  {
    const V3 cand = vadd(L.rayD, vscale(r0, L.Nrm));
    const float align = vdot(L.rayD, cand);
    if (align > bestAlign) { bestAlign = align; rdir = cand; }
  }
  if (nsol == 2) {
    const V3 cand = vadd(L.rayD, vscale(r1, L.Nrm));
    const float align = vdot(L.rayD, cand);
    if (align > bestAlign) { bestAlign = align; rdir = cand; }
  }

  float cosA2 = ex_sqrt(ex_sub(1.f, ex_mul(sinA2, sinA2)));
  if (cosA1 < 0.f) cosA2 = -cosA2;
  const float Rs = fresnel_term(n1, n2, cosA1, cosA2);
  const float Rp = fresnel_term(n1, n2, cosA2, cosA1);
  const float R = (float)exd_mul((double)ex_add(Rs, Rp), 0.5);
  const V3 rint = vscale(ex_sub(1.f, R), incI);

  /* reflection intensity of this call, raytracer.h:563-578 */
  const float prod = ex_mul(transparency, R);
  V3 rc = vscale(prod, mk(1.f, 1.f, 1.f));
  rc = vadd(rc, vscale(medA.w, mk(objB.x, objB.y, objB.z)));
  rc = vmul(L.rayI, rc);

  /* suspend (always fits: depth <= S-1) */
  ++L.top;
  Frame& f = stack[L.top];
  f.colour = L.colour; f.stage = 1.f;
  f.reflCol = rc; f.medium = (float)L.medium;
  if (significant(rc)) {
    /* raytracer.h:817-842 */
    const float perp = ex_mul(2.f, vdot(L.rayD, L.Nrm));
    const V3 rd = vunit(vsub(L.rayD, vscale(perp, L.Nrm)));
    f.reflD = rd;
    f.reflO = vadd(L.P, vscale(0.01f, rd));
  } else {
    f.reflD = mk(0.f, 0.f, 0.f); f.reflO = mk(0.f, 0.f, 0.f);
  }
  L.result = L.colour;                                   /* raytracer.h:538 */
  if (L.top < cam.S - 1) {
    L.rayD = rdir; L.rayI = rint; L.medium = target;
    L.colour = mk(0.f, 0.f, 0.f);
    set_ray_query(L, K_TRACE, L.P, rdir);
    return ACT_QUERY;
  }
  return ACT_UNWIND;                                     /* child push dropped */
}

/* Matte term complete (raytracer.h:468-484), then transparency (raytracer.h:494). */
RT_HD int after_matte(Lane& L, const SceneView& sc, bool haveMatte) {
  const float4_ objA = sc.matA[L.obj];
  const float opacity = objA.w;
  if (haveMatte) {
    V3 w = vmul(L.rayI, mk(objA.x, objA.y, objA.z));
    w = vscale(opacity, w);
    w = vmul(L.lit, w);
    L.colour = vadd(w, L.colour);
  }
  const float transparency = ex_sub(1.f, opacity);
  if (transparency > 0.f) {
    const V3 probe = vadd(vscale(0.01f, L.rayD), L.P);   /* raytracer.h:688-692 */
    L.kind = K_CONTAIN;
    L.hitIdx = -1; L.minT = 1000.f;
    query_point(L.qy, probe);
    L.ctr.containQ++;
    return ACT_QUERY;
  }
  L.result = L.colour;
  return ACT_UNWIND;
}

/* Advance the lane after its query has been answered (L.hitIdx / L.minT).
 * Returns true when the lane's current SAMPLE is finished (L.result valid);
 * otherwise L.qy / L.kind hold the next query. */
RT_HD bool advance(Lane& L, const SceneView& sc, Frame* stack, const Camera& cam) {
  int act;
  if (L.kind == K_TRACE) {
    L.ctr.rays++;
    if (L.hitIdx < 0) {
      const float4_ medA = sc.matA[L.medium];
      L.result = vmul(L.rayI, mk(medA.x, medA.y, medA.z));   /* raytracer.h:544 */
      act = ACT_UNWIND;
    } else if (!significant(L.rayI)) {
      act = ACT_UNWIND;                                      /* result left stale */
    } else {
      /* raytracer.h:171-181 for the winning sphere */
      const V3 o = query_origin(L.qy);
      const V3 d = mk(L.qy.dx, L.qy.dy, L.qy.dz);
      const float4_ g = sc.geo[L.hitIdx];
      L.obj = L.hitIdx;
      L.P = vadd(o, vscale(L.minT, d));
      L.Nrm = vunit(vsub(L.P, mk(g.x, g.y, g.z)));
      const float opacity = sc.matA[L.obj].w;
      if (opacity > 0.f) {
        L.lit = mk(0.f, 0.f, 0.f);
        L.light = 0;
        act = (sc.nl > 0) ? ACT_SHADOW : ACT_MATTE;
      } else {
        act = ACT_NOMATTE;
      }
    }
  } else if (L.kind == K_SHADOW) {
    L.ctr.rays++; L.ctr.shadow++;
    const V3 d = mk(L.qy.dx, L.qy.dy, L.qy.dz);
    bool blocked = false;
    if (L.hitIdx >= 0) {
      const V3 dist = vscale(L.minT, d);
      blocked = vdot(dist, dist) < L.gap;                  /* raytracer.h:299 */
    }
    if (!blocked) {
      /* raytracer.h:337-361; `d` is the same normalised vector calculateMatte rebuilds */
      const float4_ lp = sc.lpos[L.light], lc = sc.lcol[L.light];
      const V3 dist = vsub(mk(lp.x, lp.y, lp.z), L.P);
      const float incidence = vdot(L.Nrm, d);
      if (incidence > 0.f) {
        const float k = ex_div(incidence, vdot(dist, dist));
        L.lit = vadd(L.lit, vscale(k, mk(lc.x, lc.y, lc.z)));
      }
    }
    act = (++L.light < (int)sc.nl) ? ACT_SHADOW : ACT_MATTE;
  } else if (L.kind == K_CONTAIN) {
    /* reference loop iterations (early return at the first container) */
    L.ctr.containT += (L.hitIdx >= 0) ? (uint32_t)(L.hitIdx + 1) : sc.n;
    act = after_contain(L, sc, stack, cam);
  } else {
    return false;
  }
  if (act >= ACT_MATTE) act = after_matte(L, sc, act == ACT_MATTE);
  if (act == ACT_SHADOW) { setup_shadow(L, sc); return false; }
  if (act == ACT_UNWIND) return unwind(L, stack, cam);
  return false;
}

/* Exact test of candidate sphere i for a query of the given kind; updates the running
 * closest hit (ray kinds) or first container (containment probe). */
struct HitAcc { float minT; int hitIdx; };
RT_HD HitAcc resolve_one(HitAcc h, int kind, V3 o, V3 d, float4_ g, uint32_t i) {
  if (kind == K_CONTAIN) {
    if (h.hitIdx < 0 && contains_exact(g, o)) h.hitIdx = (int)i;
  } else {
    float t;
    if (ray_sphere_exact(g, o, d, t)) {
      if (t < h.minT) { h.minT = t; h.hitIdx = (int)i; }   /* strict: first index wins ties */
    }
  }
  return h;
}
RT_HD void resolve_candidate(Lane& L, const SceneView& sc, uint32_t i) {
  if (i >= sc.n) return;
  L.ctr.exactTests++;
  HitAcc h; h.minT = L.minT; h.hitIdx = L.hitIdx;
  h = resolve_one(h, L.kind, query_origin(L.qy), mk(L.qy.dx, L.qy.dy, L.qy.dz), sc.geo[i], i);
  L.minT = h.minT; L.hitIdx = h.hitIdx;
}

}  // namespace rtg
#endif
