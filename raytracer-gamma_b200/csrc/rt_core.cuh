/* rt_core.cuh — the trace machine of the B200 trace loop (host/device portable).
 *
 * What it computes is the reference's CPU copy of the algorithm
 * (/root/reference/raytracer_gamma/raytracer.h:81-842, pixel loop
 * main.cpp:383-453); how it computes it is new:
 *
 *  - The reference recurses through an explicit stack of 168-byte snapshots
 *    (raytraceStack.h:13-68).  Here a pixel in flight is a SLOT: a small state
 *    record whose only suspension point is "one QUERY against all N spheres".
 *    A query is a trace ray (closest hit, raytracer.h:145), a BATCH of up to four
 *    shadow rays that share their origin (one per light, raytracer.h:272,328) or
 *    a point-containment probe (raytracer.h:245).  The kernel (rt_kernels.cuh)
 *    keeps several slots per lane and, every pass, lets the warp vote for the
 *    query kind most of its lanes can serve, so each pass over the spheres runs a
 *    loop specialised for ONE kind with all lanes converged.
 *  - A suspended call keeps 14 words (colour, the pre-computed reflected ray and
 *    its intensity, the medium) instead of a 168-byte snapshot.
 *  - Every value that can change a discrete decision or is carried into the
 *    image is evaluated with single IEEE-754 binary32 operations in the
 *    reference's expression order (the ex_* helpers: __fmul_rn/__fadd_rn/
 *    __fdiv_rn/__fsqrt_rn never contract into FMAs), so the float framebuffer is
 *    bit-identical to the reference CPU render.  The FMA filter only decides
 *    which spheres CANNOT be hit; see "Filter" below for the bound.
 *
 * The file is host/device portable on purpose: tests/hostsim.cpp runs the same
 * machine slot by slot on the CPU to check it against the oracle without a GPU.
 * That build is test infrastructure; the product path is the CUDA kernel only.
 */
#ifndef RT_CORE_CUH
#define RT_CORE_CUH

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
/* out-of-line on the device: IEEE division / square root expand to ~15 instructions plus a
 * slow path each, and the O(1) shading code uses them ~50 times; one copy keeps the kernel's
 * instruction footprint (and with it the instruction-cache miss rate) small */
#define RT_HD_NI __host__ __device__ __noinline__
#else
#define RT_HD static inline
#define RT_HD_NI static
#endif

namespace rtg {

/* ---- exact (never contracted) IEEE binary32 / binary64 primitives ---------- */
#if defined(__CUDA_ARCH__)
RT_HD float ex_add(float a, float b) { return __fadd_rn(a, b); }
RT_HD float ex_sub(float a, float b) { return __fadd_rn(a, -b); }
RT_HD float ex_mul(float a, float b) { return __fmul_rn(a, b); }
RT_HD_NI float ex_div(float a, float b) { return __fdiv_rn(a, b); }
RT_HD_NI float ex_sqrt(float a) { return __fsqrt_rn(a); }
/* inline twins for the few hot spots where independent divisions / roots should overlap */
RT_HD float ex_div_i(float a, float b) { return __fdiv_rn(a, b); }
RT_HD float ex_sqrt_i(float a) { return __fsqrt_rn(a); }
RT_HD double exd_add(double a, double b) { return __dadd_rn(a, b); }
RT_HD double exd_sub(double a, double b) { return __dadd_rn(a, -b); }
RT_HD double exd_mul(double a, double b) { return __dmul_rn(a, b); }
RT_HD_NI double exd_div(double a, double b) { return __ddiv_rn(a, b); }
RT_HD_NI double exd_sqrt(double a) { return __dsqrt_rn(a); }
/* sin = (float)sqrt(1.0 - (double)(c*c)) as raytracer.h:683 computes it, for -1 < c < 1 (or NaN).
 * The library's IEEE double square root is ~290 instructions (special cases, denormals); its argument here is
 * always a normal number in (2^-24, 1], so the plain Markstein sequence — reciprocal-square-root seed (MUFU.RSQ64H),
 * two coupled Newton steps, one exact-residual correction — gives the same correctly rounded double in 14
 * (tests/fuzz_filter.cu checks it against __dsqrt_rn for EVERY float c in [0, 1)). */
RT_HD float sin_from_cos(float c) {
  const double x = __dadd_rn(1.0, -(double)__fmul_rn(c, c));
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = __dmul_rn(x, y), h = __dmul_rn(0.5, y);
  double r = __fma_rn(-g, h, 0.5);
  g = __fma_rn(g, r, g); h = __fma_rn(h, r, h);
  r = __fma_rn(-g, h, 0.5);
  g = __fma_rn(g, r, g); h = __fma_rn(h, r, h);
  const double d = __fma_rn(-g, g, x);
  g = __fma_rn(d, h, g);
  return (float)g;
}
RT_HD float fast_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
/* host build: compiled with -ffp-contract=off, so these stay separate operations */
RT_HD float ex_add(float a, float b) { return a + b; }
RT_HD float ex_sub(float a, float b) { return a - b; }
RT_HD float ex_mul(float a, float b) { return a * b; }
RT_HD float ex_div(float a, float b) { return a / b; }
RT_HD float ex_sqrt(float a) { return sqrtf(a); }
RT_HD float ex_div_i(float a, float b) { return a / b; }
RT_HD float ex_sqrt_i(float a) { return sqrtf(a); }
RT_HD double exd_add(double a, double b) { return a + b; }
RT_HD double exd_sub(double a, double b) { return a - b; }
RT_HD double exd_mul(double a, double b) { return a * b; }
RT_HD double exd_div(double a, double b) { return a / b; }
RT_HD double exd_sqrt(double a) { return sqrt(a); }
RT_HD float sin_from_cos(float c) { return (float)sqrt(1.0 - (double)(c * c)); }
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
RT_HD V3 vunit_i(V3 v) { float l = ex_div_i(1.f, ex_sqrt_i(vdot(v, v))); return vscale(l, v); }
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
  /* optional two-level form (see "Cluster filter"): cfilt | mfilt | midx are contiguous */
  const float4_* cfilt;  /* [ncPad] filter record of each cluster's bounding sphere     */
  const float4_* mfilt;  /* [nc*RT_CLUSTER] the members' filter records, cluster by cluster */
  const unsigned short* midx;  /* [nc*RT_CLUSTER] the members' sphere indices            */
  uint32_t nc, ncPad;
};

/* ---- Filter ---------------------------------------------------------------
 * For a query with origin o and direction d (|d|^2 = A) and sphere (c, r), the
 * reference's float radicand (raytracer.h:95-105) approximates
 *     D = 4 [ (d.delta)^2 - A (|delta|^2 - r^2) ],   delta = o - c.
 * The filter evaluates, with FMAs and d pre-scaled to unit length,
 *     lhs = (d'.(o-c))^2 - q            q  = |o|^2 (1-kappa)   (stored negated: nq)
 *     ch  = w - 2 o.c                    w  = |c|^2 - r^2 - kappa(|c|^2+r^2) - 2.5e-6 r - 1e-11
 * and declares a CERTAIN MISS iff lhs - ch < 0, i.e. iff
 *     (d'.delta)^2 < |delta|^2 - r^2 - kappa(|o|^2+|c|^2+r^2) - (container slack).
 * kappa = 2^-17 = 128 u covers (first-order, u = 2^-24): 17u A(|delta|^2+r^2)
 * rounding of the reference's own non-fused radicand, 4u for its rounded delta,
 * 24u(|o|^2+|c|^2) for the expanded-form d'.(o-c) squared, 10u(|o|^2+|c|^2+r^2)
 * for ch, 8u|delta|^2 for the unit scaling — about 92u(|o|^2+|c|^2)+31u r^2 in
 * total, using |delta|^2 <= 2(|o|^2+|c|^2).  Anything not a certain miss goes
 * through the reference's exact expressions, so a loose bound costs time, never
 * correctness.  The ch part depends on the origin only: the shadow rays of one hit
 * share it.  With no direction (lhs = -q) the same records answer the containment
 * probe of raytracer.h:245-270: -q - ch < 0 <=> |p-c|^2 > r^2 + slack, and the slack
 * exceeds (r+1e-6)^2 - r^2.
 * The kernel forms e = -q - ch once per origin and then fma(b, b, e) per ray: e and b are
 * finite (or e = -inf for padding records), a fused multiply-add rounds once, and a rounded
 * value carries the sign of the exact one (an exact zero gives +0), so the kernel just
 * collects SIGN BITS.  Geometry that is not finite never reaches the
 * filter (origin_filterable / dir_filterable): it is tested exactly against every
 * sphere instead.
 */
#define RT_KAPPA 7.62939453125e-06f   /* 2^-17 */

/* ---- Cluster filter (optional accelerated mode) ---------------------
 * The spheres are grouped (host side, rt_soa.h) into spatial clusters of RT_CLUSTER members.
 * Each cluster gets the filter record of a bounding sphere (C, R); a query first runs the
 * filter over the clusters and then over the members of the clusters it could not rule out.
 * Exactness needs: "the reference reports a hit of member i  =>  cluster(i) is not ruled out".
 *   (1) a reference hit means, in exact arithmetic, dist(line, c_i)^2 <= r_i^2 + eta_i with
 *       eta_i <= 64u (|o|^2 + |c_i|^2 + r_i^2)   (the 17u+4u radicand terms of "Filter", u = 2^-24);
 *       the containment probe of raytracer.h:245-270 adds 2.1e-6 r_i + 1e-12;
 *   (2) dist(line, C) <= dist(line, c_i) + D_i with D_i = |C - c_i|, so
 *       dist(line, C)^2 <= (r_i + D_i)^2 + D_i^2/32 + 33 eta_i            (2ab <= a^2/32 + 32 b^2);
 *   (3) the cluster record is the ordinary filter record of (C, R) with
 *       R^2 >= max_i (r_i + D_i)^2 + D_i^2/32 + 2^-12 (|c_i|^2 + r_i^2) + 1e-4 r_i + 1e-10,
 *       and the cluster pass uses the origin term q(1-kappa)(1-2^-12): the ordinary filter rules a
 *       sphere out only if dist^2 > R^2 (its own rounding is inside kappa), the extra 2^-12 |o|^2
 *       covers the |o|^2 share of 33 eta_i = 2112u (...) < 2^-12 (...).
 * Candidates then reach the exact expressions in cluster order, so the resolve steps break
 * ties by sphere index explicitly (first index wins, raytracer.h:166-188 / :264).
 */
#define RT_CLUSTER 8
#define RT_KAPPA2 2.44140625e-04f     /* 2^-12 */

struct OriginQ { float px, py, pz, nq; };    /* -2o, -|o|^2(1-kappa) (never -0) */
struct DirQ { float ndx, ndy, ndz, od; };    /* -d', d'.o   (d' = d/|d|)       */

RT_HD OriginQ make_origin(V3 o) {
  OriginQ O;
  O.px = -2.f * o.x; O.py = -2.f * o.y; O.pz = -2.f * o.z;
  const float oo = fast_fma(o.z, o.z, fast_fma(o.y, o.y, ex_mul(o.x, o.x)));
  O.nq = ex_sub(0.f, ex_mul(oo, 1.f - RT_KAPPA));   /* 0 - q: +0 when q is 0, so no test value is ever -0 */
  return O;
}

/* Direction part of a ray query.  Returns false when d == 0: the reference then
 * computes a = b = 0, radicand = 0, u = 0/0 = NaN for every sphere
 * (raytracer.h:98-116), i.e. a guaranteed miss — such a ray needs no sphere loop. */
RT_HD bool make_dir(DirQ& D, V3 o, V3 d) {
  const float A = vdot(d, d);
  if (!(A > 0.f) && !(A != A)) {   /* A == 0 (not NaN) */
    D.ndx = D.ndy = D.ndz = 0.f; D.od = 0.f;
    return false;
  }
  const float s = ex_div(1.f, ex_sqrt(A));
  const float ux = ex_mul(d.x, s), uy = ex_mul(d.y, s), uz = ex_mul(d.z, s);
  D.ndx = -ux; D.ndy = -uy; D.ndz = -uz;
  D.od = fast_fma(uz, o.z, fast_fma(uy, o.y, ex_mul(ux, o.x)));
  return true;
}

/* The filter's direction only has to be unit length to within the bound's "unit scaling" share
 * (see Filter): two cheaper set-ups than make_dir's IEEE division and square root.
 *  - make_dir_fast: d' = d * rsqrt(d.d) with the hardware approximation (MUFU.RSQ, relative error
 *    <= 2^-22.9 by the PTX ISA): |d'|^2 = 1 +- 12u (2 x 4.3u for the approximation, 2u for the rounded
 *    d.d, 2u for the three products), i.e. (d'.delta)^2 is off by <= 12u |delta|^2 <= 24u(|o|^2+|c|^2)
 *    where make_dir's share was 16u: the total first-order bound grows from ~92u to ~100u, kappa = 128u.
 *  - make_dir_unit: d is already the reference's own normalised vector (vec.h:41: |d|^2 = 1 +- 3u),
 *    used as it is.
 * A direction whose squared length is zero, subnormal, infinite or NaN is never filtered: zero is the
 * reference's certain miss (returns false), everything else goes through the exact test (D.od = NaN
 * makes dir_filterable fail). */
RT_HD float rsqrt_fast(float a) {
#if defined(__CUDA_ARCH__)
  float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r;
#else
  return 1.f / sqrtf(a);
#endif
}
RT_HD bool make_dir_fast(DirQ& D, V3 o, V3 d) {
  const float A = vdot(d, d);
  if (A == 0.f) { D.ndx = D.ndy = D.ndz = 0.f; D.od = 0.f; return false; }
  const float s = rsqrt_fast(A);
  const float ux = ex_mul(d.x, s), uy = ex_mul(d.y, s), uz = ex_mul(d.z, s);
  D.ndx = -ux; D.ndy = -uy; D.ndz = -uz;
  D.od = fast_fma(uz, o.z, fast_fma(uy, o.y, ex_mul(ux, o.x)));
  if (!(A >= 1.17549435e-38f && A <= 3.0e38f)) D.od = HUGE_VALF;   /* not filterable: exact test against every sphere */
  return true;
}
RT_HD void make_dir_unit(DirQ& D, V3 o, V3 d) {
  D.ndx = -d.x; D.ndy = -d.y; D.ndz = -d.z;
  D.od = fast_fma(d.z, o.z, fast_fma(d.y, o.y, ex_mul(d.x, o.x)));
  /* vnorm of a denormal / overflowing vector is not unit: such a ray is tested exactly (8u = the bound's share) */
  const float A = fast_fma(d.z, d.z, fast_fma(d.y, d.y, ex_mul(d.x, d.x)));
  if (!(fabsf(A - 1.f) <= 4.76837158e-07f)) D.od = HUGE_VALF;
}

/* origin term of the cluster pass: q(1-kappa)(1-2^-12), see "Cluster filter" */
RT_HD OriginQ cluster_origin(const OriginQ& O) {
  OriginQ C = O;
  C.nq = ex_mul(O.nq, 1.f - RT_KAPPA2);
  return C;
}

RT_HD float filter_ch(const OriginQ& O, float4_ s) {
  float ch = fast_fma(O.px, s.x, s.w);
  ch = fast_fma(O.py, s.y, ch);
  return fast_fma(O.pz, s.z, ch);
}
/* Ray test: negative <=> certain miss.  e = -q - ch is shared by all rays of one origin. */
RT_HD float filter_ray(const OriginQ& O, const DirQ& D, float ch, float4_ s) {
  float bq = fast_fma(D.ndx, s.x, D.od);
  bq = fast_fma(D.ndy, s.y, bq);
  bq = fast_fma(D.ndz, s.z, bq);
  const float e = ex_sub(O.nq, ch);
  return fast_fma(bq, bq, e);
}
/* Containment test: negative <=> certainly outside. */
RT_HD float filter_point(const OriginQ& O, float ch) { return ex_sub(O.nq, ch); }

RT_HD bool finite_f(float x) { return fabsf(x) <= 3.402823466e+38f; }
RT_HD bool origin_filterable(const OriginQ& O) {
  return finite_f(O.px) && finite_f(O.py) && finite_f(O.pz) && finite_f(O.nq);
}
RT_HD bool dir_filterable(const DirQ& D) {
  return finite_f(D.ndx) && finite_f(D.ndy) && finite_f(D.ndz) && finite_f(D.od);
}

/* ---- exact per-candidate tests --------------------------------------------- */
/* raytracer.h:81-141.  The reference forms both roots u0 = (-b+root)/denom and
 * u1 = (-b-root)/denom and keeps the smallest one in (1e-5, 10000).  Rounding is
 * monotonic and denom = 2 d.d > 0, so u1 <= u0 always: when u1 > 1e-5 it is the
 * answer (or, if u1 >= 10000, so is u0 and there is none) and u0 need not be
 * divided out; u0 is only formed when u1 fails the lower bound (or is NaN).
 * Same decisions, same t, one IEEE division less on the common path. */
RT_HD_NI float ray_sphere_t(float4_ g, V3 o, V3 d) {   /* t of the hit, or -1 when there is none */
  const V3 disp = vsub(o, mk(g.x, g.y, g.z));
  const float a = vdot(d, d);
  const float b = ex_mul(2.0f, vdot(d, disp));
  const float c = ex_sub(vdot(disp, disp), ex_mul(g.w, g.w));
  const float radicand = ex_sub(ex_mul(b, b), ex_mul(ex_mul(4.0f, a), c));
  if (!(radicand >= 0.0f)) return -1.f;
  const float root = ex_sqrt(radicand);
  const float denom = ex_mul(2.0f, a);
  const float u1 = ex_div(ex_sub(-b, root), denom);
  if (u1 > 1.0e-5f) return (u1 < 10000.f) ? u1 : -1.f;
  const float u0 = ex_div(ex_add(-b, root), denom);
  return (u0 > 1.0e-5f && u0 < 10000.f) ? u0 : -1.f;
}
RT_HD bool ray_sphere_exact(float4_ g, V3 o, V3 d, float& t) {
  const float r = ray_sphere_t(g, o, d);
  if (r > 0.f) { t = r; return true; }
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
#if defined(__CUDACC__)
#define RT_FRAME_STRIDE 32     /* distance in frames between consecutive levels of one stack: the kernel keeps the 32 lanes' frames of a level side by side */
#else
#define RT_FRAME_STRIDE 1
#endif
#define RT_SHADOW_BATCH 4
/* A suspended call: 13 words laid out as one 64-byte chunk = two 32-byte sectors.  Every pop reads the first
 * sector (colour, tag, reflection intensity: see unwind); only a pop that launches the reflected child reads the
 * second (the pre-computed ray). */
struct
#if defined(__CUDACC__)
__align__(16)
#endif
Frame {
  V3 colour;              /* cur.colour at suspension                                  */
  int tag;                /* stage << 16 | medium: stage 1 waits for the refracted child, 2 for the reflected one */
  V3 reflCol;             /* reflection intensity (raytracer.h:563-578)                */
  int pad0;
  V3 reflO;               /* pre-computed reflected ray (raytracer.h:817-842)           */
  int pad1;
  V3 reflD;
  int pad2;
};
RT_HD int frame_tag(int stage, int medium) { return (stage << 16) | medium; }
RT_HD int frame_stage(const Frame& f) { return f.tag >> 16; }
RT_HD int frame_medium(const Frame& f) { return f.tag & 0xFFFF; }
/* Frame traffic: 16-byte vector accesses (ordinary caching: bypassing L1 with ld.cg / st.cg measured 0.5 % slower). */
#if defined(__CUDA_ARCH__)
RT_HD void frame_load_hot(const Frame* f, V3& colour, V3& reflCol, int& tag) {
  const float4 a = *reinterpret_cast<const float4*>(f), b = *(reinterpret_cast<const float4*>(f) + 1);
  colour = mk(a.x, a.y, a.z); tag = __float_as_int(a.w); reflCol = mk(b.x, b.y, b.z);
}
RT_HD void frame_load_ray(const Frame* f, V3& reflO, V3& reflD) {
  const float4 c = *(reinterpret_cast<const float4*>(f) + 2), d = *(reinterpret_cast<const float4*>(f) + 3);
  reflO = mk(c.x, c.y, c.z); reflD = mk(d.x, d.y, d.z);
}
RT_HD void frame_store_head(Frame* f, V3 colour, int tag) {
  *reinterpret_cast<float4*>(f) = make_float4(colour.x, colour.y, colour.z, __int_as_float(tag));
}
RT_HD void frame_store(Frame* f, const Frame& v) {
  float4* q = reinterpret_cast<float4*>(f);
  q[0] = make_float4(v.colour.x, v.colour.y, v.colour.z, __int_as_float(v.tag));
  q[1] = make_float4(v.reflCol.x, v.reflCol.y, v.reflCol.z, 0.f);
  q[2] = make_float4(v.reflO.x, v.reflO.y, v.reflO.z, 0.f);
  q[3] = make_float4(v.reflD.x, v.reflD.y, v.reflD.z, 0.f);
}
#else
RT_HD void frame_load_hot(const Frame* f, V3& colour, V3& reflCol, int& tag) { colour = f->colour; reflCol = f->reflCol; tag = f->tag; }
RT_HD void frame_load_ray(const Frame* f, V3& reflO, V3& reflD) { reflO = f->reflO; reflD = f->reflD; }
RT_HD void frame_store_head(Frame* f, V3 colour, int tag) { f->colour = colour; f->tag = tag; }
RT_HD void frame_store(Frame* f, const Frame& v) { *f = v; }
#endif

struct Counters {         /* per-lane tallies, reduced per block at the end */
  uint32_t rays, shadow, containQ, containT, exactTests, samples;
  uint32_t nullRays;      /* rays with d == 0: answered without running the sphere loop */
};

enum { K_NULL = 0, K_TRACE = 1, K_SHADOW = 2, K_CONTAIN = 3 };

/* ---- a pixel in flight --------------------------------------------------------- */
struct Slot {
  int kind;               /* the pending query */
  uint32_t pixel;         /* index of the sample's result record, 0xFFFFFFFF = free slot */
  int top;                /* frames on the stack = depth of the current call */
  int medium;             /* sphere index of the medium the ray travels in, n = ambient */
  int obj;                /* struck sphere */
  int light;              /* first light of the current shadow batch */
  int ndirs;              /* directions in the pending query (1 for a trace ray) */
  int count;              /* queries of this sample answered so far, saturating at RT_COUNT_MAX (work order, see WorkMap) */
  int first;              /* 1: the sample belongs to its tile's FIRST group, whose chain lengths order the tile's other groups */
  V3 result;              /* "colourSum" of raytracer.h:425 */
  V3 colour, rayD, rayI;  /* the call being evaluated ("currSnapshot") */
  V3 P, Nrm;              /* its hit */
  V3 lit;                 /* matte accumulation, raytracer.h:325 */
  /* geometry of the pending query: a trace ray is (qo, rayD); a containment probe is the
   * point qo; a shadow batch starts at P towards lights light..light+ndirs-1 — its
   * directions live only for the pass that serves it (ShadowGeo), not in the record */
  V3 qo;
  /* answer, filled by the sphere pass */
  float minT; int hitIdx; /* closest hit (trace) / first container (probe) */
  uint32_t blocked;       /* shadow batch: bit k = light `light+k` is occluded */
};
/* The suspended calls of a slot live beside it: Frame stack[RT_MAX_STACK] (only the slot
 * record is small enough to be copied into registers while it is advanced). */

/* ---- slot records: how a slot is kept between two passes (21 words) ------------------------
 * word 0      pixel (index of the sample's result record, 0xFFFFFFFF = free slot)
 * word 1      kind | ndirs << 2 | (top + 1) << 5 | light << 10 (15 bits) | count << 25 (6 bits) | first << 31
 * word 2      medium | obj << 16
 * words 3..   result, rayD, rayI, P, Nrm, lit|colour (3 words each).  While a TRACE query is pending the
 *             hit fields are dead and the P words hold the ray's origin; the probe point of a CONTAIN
 *             query is P + 0.01 rayD (raytracer.h:688-692), rebuilt when the record is read.  The call's
 *             colour is zero from its start until the matte term is added (raytracer.h:468-484), i.e.
 *             whenever a TRACE or SHADOW query is pending, and the matte sum `lit` is dead once it has
 *             been added: the two share their words (lit while SHADOW is pending, colour while CONTAIN is).
 * The answer of the query (minT / hitIdx / blocked) is not part of the record: the pass hands it over in
 * registers.  STRIDE = distance in words between consecutive record words (1, or the CTA size when the
 * records of a CTA are interleaved in shared memory). */
#define RT_SLOT_WORDS 21
#define RT_COUNT_MAX 63
enum { W_PIXEL = 0, W_HDR = 1, W_MEDOBJ = 2, W_RESULT = 3, W_RAYD = 6, W_RAYI = 9, W_P = 12, W_NRM = 15, W_LITCOL = 18 };
#if defined(__CUDA_ARCH__)
RT_HD uint32_t f2u(float f) { return __float_as_uint(f); }
RT_HD float u2f(uint32_t u) { return __uint_as_float(u); }
#else
RT_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
RT_HD float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#endif
template <int STRIDE> RT_HD V3 rec_ldv(const uint32_t* r, int wd) {
  return mk(u2f(r[wd * STRIDE]), u2f(r[(wd + 1) * STRIDE]), u2f(r[(wd + 2) * STRIDE]));
}
template <int STRIDE> RT_HD void rec_stv(uint32_t* r, int wd, V3 v) {
  r[wd * STRIDE] = f2u(v.x); r[(wd + 1) * STRIDE] = f2u(v.y); r[(wd + 2) * STRIDE] = f2u(v.z);
}
template <int STRIDE> RT_HD void slot_pack(uint32_t* r, const Slot& s) {
  r[W_PIXEL * STRIDE] = s.pixel;
  r[W_HDR * STRIDE] = (uint32_t)s.kind | ((uint32_t)s.ndirs << 2) | ((uint32_t)(s.top + 1) << 5) | ((uint32_t)s.light << 10) |
                      ((uint32_t)s.count << 25) | ((uint32_t)s.first << 31);
  r[W_MEDOBJ * STRIDE] = ((uint32_t)s.medium & 0xFFFFu) | ((uint32_t)s.obj << 16);
  rec_stv<STRIDE>(r, W_RESULT, s.result);
  rec_stv<STRIDE>(r, W_RAYD, s.rayD); rec_stv<STRIDE>(r, W_RAYI, s.rayI);
  rec_stv<STRIDE>(r, W_P, (s.kind == K_TRACE) ? s.qo : s.P);
  rec_stv<STRIDE>(r, W_NRM, s.Nrm);
  rec_stv<STRIDE>(r, W_LITCOL, (s.kind == K_CONTAIN) ? s.colour : s.lit);
}
template <int STRIDE> RT_HD void slot_unpack(const uint32_t* r, Slot& s) {
  s.pixel = r[W_PIXEL * STRIDE];
  const uint32_t h = r[W_HDR * STRIDE], mo = r[W_MEDOBJ * STRIDE];
  s.kind = (int)(h & 3u); s.ndirs = (int)((h >> 2) & 7u); s.top = (int)((h >> 5) & 31u) - 1; s.light = (int)((h >> 10) & 0x7FFFu);
  s.count = (int)((h >> 25) & 63u); s.first = (int)(h >> 31);
  s.medium = (int)(mo & 0xFFFFu); s.obj = (int)(mo >> 16);
  s.result = rec_ldv<STRIDE>(r, W_RESULT);
  s.rayD = rec_ldv<STRIDE>(r, W_RAYD); s.rayI = rec_ldv<STRIDE>(r, W_RAYI);
  s.P = rec_ldv<STRIDE>(r, W_P); s.Nrm = rec_ldv<STRIDE>(r, W_NRM);
  const V3 lc = rec_ldv<STRIDE>(r, W_LITCOL), zero = mk(0.f, 0.f, 0.f);
  s.colour = (s.kind == K_CONTAIN) ? lc : zero;
  s.lit = (s.kind == K_CONTAIN) ? zero : lc;
  s.qo = (s.kind == K_CONTAIN) ? vadd(vscale(0.01f, s.rayD), s.P) : s.P;
}

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


/* ---- work items -----------------------------------------------------------------
 * A work item is one sample of one pixel; 32 consecutive items form a GROUP (tile, k):
 *   spp > 1:  tile = 8x4 pixels, k = sample index (K = spp groups per tile);
 *   spp == 1: tile = 16x8 pixels, k = one of its four stride-2 sub-lattices (K = 4) —
 * either way the groups of a tile see almost the same geometry, so the chain lengths measured on group 0
 * predict the others: the kernel hands out every tile's group 0 first and then the remaining groups
 * deepest tiles first (longest-processing-time order; the last samples of a launch are then short ones
 * and the launch does not end on a few warps working off 100-query chains).  Rows are the context's LOCAL
 * rows (strip r of the frame is ours iff r % stripStride == stripFirst). */
struct WorkMap {
  uint32_t W, localRows, stripRows, stripFirst, stripStride;
  uint32_t spp, nIter, K, lattice, tilesX, tilesY, nTiles;
};
RT_HD WorkMap make_workmap(uint32_t W, uint32_t localRows, uint32_t stripRows, uint32_t stripFirst, uint32_t stripStride,
                           uint32_t spp, uint32_t nIter) {
  WorkMap m;
  m.W = W; m.localRows = localRows; m.stripRows = stripRows; m.stripFirst = stripFirst; m.stripStride = stripStride;
  m.spp = spp; m.nIter = nIter;
  m.lattice = (spp == 1u) ? 1u : 0u;
  m.K = m.lattice ? 4u : spp;
  m.tilesX = m.lattice ? (W + 15u) / 16u : (W + 7u) / 8u;
  m.tilesY = m.lattice ? (localRows + 7u) / 8u : (localRows + 3u) / 4u;
  m.nTiles = m.tilesX * m.tilesY;
  return m;
}
/* item `within` (0..31) of group (tile, k): frame pixel (gx, gy), index of the result record, sample (si, sj).
 * false = the item lies outside the frame. */
RT_HD bool work_item(const WorkMap& m, uint32_t tile, uint32_t k, uint32_t within, uint32_t& gx, uint32_t& gy,
                     uint32_t& dst, int& si, int& sj) {
  const uint32_t ty = tile / m.tilesX, tx = tile - ty * m.tilesX;
  uint32_t x, y;
  if (m.lattice) { x = tx * 16u + 2u * (within & 7u) + (k & 1u); y = ty * 8u + 2u * (within >> 3) + (k >> 1); }
  else           { x = tx * 8u + (within & 7u);                 y = ty * 4u + (within >> 3); }
  if (x >= m.W || y >= m.localRows) return false;
  const uint32_t strip = y / m.stripRows;
  gx = x;
  gy = (strip * m.stripStride + m.stripFirst) * m.stripRows + (y - strip * m.stripRows);
  if (m.lattice) { dst = y * m.W + x; si = 0; sj = 0; }
  else {
    dst = (y * m.W + x) * m.spp + k;
    si = (int)(k / m.nIter);
    sj = (int)(k - (uint32_t)si * m.nIter);
  }
  return true;
}
/* the tile a result record belongs to */
RT_HD uint32_t tile_of_dst(const WorkMap& m, uint32_t dst) {
  const uint32_t pix = m.lattice ? dst : dst / m.spp;
  const uint32_t y = pix / m.W, x = pix - y * m.W;
  return m.lattice ? (y >> 3) * m.tilesX + (x >> 4) : (y >> 2) * m.tilesX + (x >> 3);
}
/* work order: a tile whose first group shows a chain of RT_DEEP_AT queries is handed out ahead of the others */
#define RT_DEEP_AT 24

/* The caller has already made `d` the current call's direction (s.rayD). */
RT_HD void set_trace_query(Slot& s, Counters& ctr, V3 o, V3 d) {
  s.kind = K_TRACE;
  s.ndirs = 1;
  s.qo = o;
  s.minT = 1000.f; s.hitIdx = -1;
  if (vdot(d, d) == 0.f) ctr.nullRays++;   /* zero direction: certain miss (see make_dir) */
}

/* Start sample (si, sj) of pixel (gx, gy) of the full frame (main.cpp:409-441): the slot
 * is one SAMPLE in flight.  Samples of a pixel are independent calls of rayTrace (main.cpp:439);
 * they are traced as separate work items and summed afterwards in the reference's order
 * (combine_samples / main.cpp:443-447), which shortens every dependency chain by the sample
 * count.  `pixel` is the index of the sample's result record. */
RT_HD void start_task(Slot& s, Counters& ctr, const Camera& cam, uint32_t gx, uint32_t gy,
                      uint32_t pixel, int si, int sj) {
  s.pixel = pixel;
  const float pxw = ex_mul(ex_sub((float)gx, ex_mul((float)cam.W, 0.5f)), cam.stepX);
  const float pyw = ex_mul(ex_sub(ex_mul((float)cam.H, 0.5f), (float)gy), cam.stepY);
  const float x = ex_mul(ex_add(pxw, ex_mul((float)sj, cam.aliasStep)), cam.aspect);
  const float y = ex_add(pyw, ex_mul((float)si, cam.aliasStep));
  const V3 d = vunit(mk(x, y, cam.zoom));
  s.rayD = d; s.rayI = mk(1.f, 1.f, 1.f); s.colour = mk(0.f, 0.f, 0.f);
  s.result = mk(0.f, 0.f, 0.f);
  s.medium = cam.ambient;                  /* every sample starts in the ambient medium, main.cpp:439 */
  s.top = -1;
  s.count = 0; s.first = 0;
  ctr.samples++;
  set_trace_query(s, ctr, mk(0.f, 0.f, 0.f), d);
}

/* The sample's contribution to its pixel: vsmul(currentSample, kSamplesTotinv, ...) main.cpp:443 */
RT_HD V3 sample_value(const Slot& s, const Camera& cam) { return vscale(cam.inv, s.result); }

/* Sum the nIter x nIter scaled samples of one pixel in the reference's loop order
 * (i outer, j inner; main.cpp:430-447). */
RT_HD V3 combine_samples(const V3* samples, int count) {
  V3 acc = mk(0.f, 0.f, 0.f);
  for (int k = 0; k < count; ++k) acc = vadd(acc, samples[k]);
  return acc;
}

/* Next shadow batch of the current hit: lights light .. light+ndirs-1. */
RT_HD void setup_shadow_batch(Slot& s, const SceneView& sc) {
  int nb = (int)sc.nl - s.light;
  if (nb > RT_SHADOW_BATCH) nb = RT_SHADOW_BATCH;
  s.kind = K_SHADOW;
  s.ndirs = nb;
  s.blocked = 0u;
}

/* The rays of a shadow batch (raytracer.h:279-286): from P towards each light, with the
 * squared distance to it.  Rebuilt by the pass that serves the batch. */
struct ShadowGeo {
  V3 d[RT_SHADOW_BATCH];        /* unit directions towards the lights of the batch                              */
  float gap[RT_SHADOW_BATCH];   /* squared distances to them (also calculateMatte's divisor, raytracer.h:349)   */
  float inc[RT_SHADOW_BATCH];   /* incidence N.d (raytracer.h:341-345), formed where the directions are         */
};
RT_HD void shadow_geo(const Slot& s, const SceneView& sc, ShadowGeo& g) {
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int k = 0; k < RT_SHADOW_BATCH; ++k) {
    if (k < s.ndirs) {
      const float4_ lp = sc.lpos[s.light + k];
      const V3 dir = vsub(mk(lp.x, lp.y, lp.z), s.P);
      g.gap[k] = vdot(dir, dir);
      g.d[k] = vunit_i(dir);       /* independent normalisations: inline so they overlap */
      g.inc[k] = vdot(s.Nrm, g.d[k]);
    } else {
      g.gap[k] = 0.f; g.d[k] = mk(0.f, 0.f, 0.f); g.inc[k] = 0.f;
    }
  }
}

/* Pop suspended calls until one launches a child ray or the stack is empty
 * (raytracer.h:552-628).  Returns true when the sample is finished. */
RT_HD bool unwind(Slot& s, Frame* stack, Counters& ctr, const Camera& cam) {
  if (s.top < 0) return true;
  /* The frames live in global memory (an L2 round trip each) and the frames of consecutive levels are independent
   * loads: the next level's words are requested while the current level is being combined, so popping k levels
   * costs one memory latency instead of k (ncu, 256 spheres: 6.5 % of all warp samples waited here). */
  V3 col, rc, ro, rd;
  int tag;
  frame_load_hot(&stack[s.top * RT_FRAME_STRIDE], col, rc, tag);
  for (;;) {
    const int lvl = s.top;
    V3 ncol = mk(0.f, 0.f, 0.f), nrc = ncol;
    int ntag = 0;
    if (lvl > 0) frame_load_hot(&stack[(lvl - 1) * RT_FRAME_STRIDE], ncol, nrc, ntag);
    --s.top;
    s.colour = vadd(s.result, col);
    if ((tag >> 16) == 1 && significant(rc)) {
      /* re-push as stage 2 (always fits: the slot was just vacated) */
      ++s.top;
      const int medium = tag & 0xFFFF;
      frame_store_head(&stack[lvl * RT_FRAME_STRIDE], s.colour, frame_tag(2, medium));
      s.result = s.colour;
      if (s.top < cam.S - 1) {
        /* reflected child, raytracer.h:602-611 */
        frame_load_ray(&stack[lvl * RT_FRAME_STRIDE], ro, rd);
        s.rayD = rd; s.rayI = rc; s.medium = medium;
        s.colour = mk(0.f, 0.f, 0.f);
        set_trace_query(s, ctr, ro, rd);
        return false;
      }
      /* child push dropped (raytraceStack.h:52): the frame just written is popped again as stage 2 */
      col = s.colour; tag = frame_tag(2, medium);
      continue;
    }
    s.result = s.colour;
    if (lvl == 0) return true;
    col = ncol; rc = nrc; tag = ntag;
  }
}

/* raytracer.h:370-403 */
RT_HD_NI float fresnel_term(float n1, float n2, float cosA1, float cosA2) {
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
 * matte/transparency step, unwinding) exists once, so the code stays small. */
enum { ACT_QUERY = 0, ACT_UNWIND = 1, ACT_SHADOW = 2, ACT_MATTE = 3, ACT_NOMATTE = 4 };

/* After the containment probe: refraction (raytracer.h:642-815), suspend the
 * call (raytracer.h:516-522), launch the refracted child (raytracer.h:524-533). */
RT_HD int after_contain(Slot& s, Frame* stack, Counters& ctr, const SceneView& sc, const Camera& cam) {
  const int target = (s.hitIdx >= 0) ? s.hitIdx : (int)sc.n;
  const float4_ objA = sc.matA[s.obj], objB = sc.matB[s.obj];
  const float4_ medA = sc.matA[s.medium], medB = sc.matB[s.medium];
  const float4_ tgtB = sc.matB[target];
  const float opacity = objA.w;
  const float transparency = ex_sub(1.f, opacity);
  const V3 incI = vscale(transparency, s.rayI);          /* raytracer.h:498 */

  float cosA1 = vdot(s.rayD, s.Nrm);
  float sinA1 = 0.f;
  if (cosA1 <= -1.0f) { cosA1 = -1.f; }
  else if (cosA1 >= 1.f) { cosA1 = 1.f; }
  else { sinA1 = sin_from_cos(cosA1); }                  /* :683, double */

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
    const V3 cand = vadd(s.rayD, vscale(r0, s.Nrm));
    const float align = vdot(s.rayD, cand);
    if (align > bestAlign) { bestAlign = align; rdir = cand; }
  }
  if (nsol == 2) {
    const V3 cand = vadd(s.rayD, vscale(r1, s.Nrm));
    const float align = vdot(s.rayD, cand);
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
  rc = vmul(s.rayI, rc);

  /* suspend (always fits: depth <= S-1): the frame is built in registers and stored as one 64-byte chunk */
  ++s.top;
  Frame f;
  f.colour = s.colour; f.tag = frame_tag(1, s.medium);
  f.reflCol = rc;
  f.pad0 = f.pad1 = f.pad2 = 0;
  if (significant(rc)) {
    /* raytracer.h:817-842 */
    const float perp = ex_mul(2.f, vdot(s.rayD, s.Nrm));
    const V3 rd = vunit(vsub(s.rayD, vscale(perp, s.Nrm)));
    f.reflD = rd;
    f.reflO = vadd(s.P, vscale(0.01f, rd));
  } else {
    f.reflD = mk(0.f, 0.f, 0.f); f.reflO = mk(0.f, 0.f, 0.f);
  }
  frame_store(&stack[s.top * RT_FRAME_STRIDE], f);
  s.result = s.colour;                                   /* raytracer.h:538 */
  if (s.top < cam.S - 1) {
    s.rayD = rdir; s.rayI = rint; s.medium = target;
    s.colour = mk(0.f, 0.f, 0.f);
    set_trace_query(s, ctr, s.P, rdir);
    return ACT_QUERY;
  }
  return ACT_UNWIND;                                     /* child push dropped */
}

/* Matte term complete (raytracer.h:468-484), then transparency (raytracer.h:494). */
RT_HD int after_matte(Slot& s, Counters& ctr, const SceneView& sc, bool haveMatte) {
  const float4_ objA = sc.matA[s.obj];
  const float opacity = objA.w;
  if (haveMatte) {
    V3 w = vmul(s.rayI, mk(objA.x, objA.y, objA.z));
    w = vscale(opacity, w);
    w = vmul(s.lit, w);
    s.colour = vadd(w, s.colour);
  }
  const float transparency = ex_sub(1.f, opacity);
  if (transparency > 0.f) {
    s.kind = K_CONTAIN;
    s.ndirs = 0;
    s.qo = vadd(vscale(0.01f, s.rayD), s.P);             /* raytracer.h:688-692 */
    s.hitIdx = -1; s.minT = 1000.f;
    ctr.containQ++;
    return ACT_QUERY;
  }
  s.result = s.colour;
  return ACT_UNWIND;
}

/* Advance the slot after its query has been answered (hitIdx/minT or blocked).
 * Returns true when the slot's current SAMPLE is finished (result valid);
 * otherwise the slot holds its next query. */
RT_HD bool advance(Slot& s, Frame* stack, Counters& ctr, const SceneView& sc, const Camera& cam,
                   const ShadowGeo* sg) {
  int act;
  if (s.count < RT_COUNT_MAX) ++s.count;
  if (s.kind == K_TRACE) {
    ctr.rays++;
    if (s.hitIdx < 0) {
      const float4_ medA = sc.matA[s.medium];
      s.result = vmul(s.rayI, mk(medA.x, medA.y, medA.z));   /* raytracer.h:544 */
      act = ACT_UNWIND;
    } else if (!significant(s.rayI)) {
      act = ACT_UNWIND;                                      /* result left stale */
    } else {
      /* raytracer.h:171-181 for the winning sphere */
      const float4_ g = sc.geo[s.hitIdx];
      s.obj = s.hitIdx;
      s.P = vadd(s.qo, vscale(s.minT, s.rayD));
      s.Nrm = vunit(vsub(s.P, mk(g.x, g.y, g.z)));
      const float opacity = sc.matA[s.obj].w;
      if (opacity > 0.f) {
        s.lit = mk(0.f, 0.f, 0.f);
        s.light = 0;
        act = (sc.nl > 0) ? ACT_SHADOW : ACT_MATTE;
      } else {
        act = ACT_NOMATTE;
      }
    }
  } else if (s.kind == K_SHADOW) {
    /* raytracer.h:328-363 for the lights of this batch, in index order */
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int k = 0; k < RT_SHADOW_BATCH; ++k) {
      if (k < s.ndirs) {
        ctr.rays++; ctr.shadow++;
        if (!((s.blocked >> k) & 1u)) {
          /* the pass that served the batch formed d = unit(light - P), |light - P|^2 and N.d: the vectors
           * calculateMatte rebuilds (raytracer.h:336-349) are the same expressions on the same operands */
          const float incidence = sg->inc[k];
          if (incidence > 0.f) {
            const float4_ lc = sc.lcol[s.light + k];
            const float kk = ex_div_i(incidence, sg->gap[k]);
            s.lit = vadd(s.lit, vscale(kk, mk(lc.x, lc.y, lc.z)));
          }
        }
      }
    }
    s.light += s.ndirs;
    act = (s.light < (int)sc.nl) ? ACT_SHADOW : ACT_MATTE;
  } else if (s.kind == K_CONTAIN) {
    /* reference loop iterations (early return at the first container) */
    ctr.containT += (s.hitIdx >= 0) ? (uint32_t)(s.hitIdx + 1) : sc.n;
    act = after_contain(s, stack, ctr, sc, cam);
  } else {
    return false;
  }
  if (act >= ACT_MATTE) act = after_matte(s, ctr, sc, act == ACT_MATTE);
  if (act == ACT_SHADOW) { setup_shadow_batch(s, sc); return false; }
  if (act == ACT_UNWIND) return unwind(s, stack, ctr, cam);
  return false;
}

/* ---- exact resolution of one candidate ---------------------------------------- */
/* Trace ray: running closest hit (raytracer.h:166-188; strict <: first index wins ties). */
RT_HD void resolve_trace(float& minT, int& hitIdx, V3 o, V3 d, float4_ g, uint32_t i) {
  float t;
  if (ray_sphere_exact(g, o, d, t)) {
    if (t < minT) { minT = t; hitIdx = (int)i; }
  }
}
/* Shadow ray towards a light at squared distance `gap`: occluded iff the closest hit with
 * t < 1000 has |t d|^2 < gap (raytracer.h:291-304).  |t d|^2 is monotone in t, so "some hit
 * with t < 1000 and |t d|^2 < gap" is the same predicate (SURVEY.md §8a note 7). */
RT_HD bool resolve_shadow(V3 o, V3 d, float gap, float4_ g) {
  float t;
  if (ray_sphere_exact(g, o, d, t) && t < 1000.f) {
    const V3 dist = vscale(t, d);
    return vdot(dist, dist) < gap;
  }
  return false;
}
/* Containment probe: first containing sphere in index order (raytracer.h:255-267). */
RT_HD void resolve_contain(int& hitIdx, V3 p, float4_ g, uint32_t i) {
  if (hitIdx < 0 && contains_exact(g, p)) hitIdx = (int)i;
}

}  // namespace rtg
#endif
