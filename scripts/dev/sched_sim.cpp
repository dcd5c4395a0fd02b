/* Development aid: discrete-event model of trace_kernel's scheduling (slots, vote, passes, drain) replaying the REAL
 * per-sample query sequences of a frame (computed with rt_core.cuh, cached in a file).  Used to try work-order / vote /
 * drain policies without a GPU.
 *   g++ -O2 -fopenmp -ffp-contract=off -mfma -Iinclude -Iraytracer-gamma_b200/csrc scripts/dev/sched_sim.cpp raytracer-gamma_b200/host/rt_scene.c
 *   sched_sim gen  <file> n W H alias S stripRows stripFirst stripStride
 *   sched_sim run  <file> [key=value ...]                                                                             */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <algorithm>
#include <queue>
#include <string>
#include <map>
#include "rt_core.cuh"
#include "rt_soa.h"
#include "rt_scene.h"
using namespace rtg;

struct Frame_ { uint32_t n, W, H, spp, S, localRows, tilesX, tilesY; };
static std::vector<uint64_t> gOff;      /* per work item: offset into gSeq (work item = tile-major index as in work_to_task) */
static std::vector<uint8_t> gSeq;       /* kinds: 1 trace, 2|nd<<2 shadow, 3 contain */
static Frame_ gF;

static void gen(const char* path, unsigned n, unsigned W, unsigned H, float alias, int S, unsigned stripRows, unsigned stripFirst, unsigned stripStride, int lattice) {
  std::vector<rt_sphere> sph(n); std::vector<rt_light> lg(4);
  rt_scene_synth(n, 4, 0, sph.data(), lg.data());
  std::vector<float4_> h; SceneLayout lay;
  build_scene_soa(sph.data(), n, lg.data(), 4, h, lay);
  const SceneView sc = scene_view(h.data(), lay);
  const Camera cam = make_camera(W, H, -4.f, alias, S, (int)sc.n);
  const uint32_t nStrips = (H + stripRows - 1) / stripRows;
  uint32_t localRows = 0;
  for (uint32_t s = stripFirst; s < nStrips; s += stripStride) localRows += std::min(stripRows, H - s * stripRows);
  /* lattice (1 spp only): a group is one of the four stride-2 sub-lattices of a 16x8 super-tile instead of one sample of an 8x4 tile */
  const uint32_t spp = lattice ? 4 : cam.nIter * cam.nIter, tilesX = lattice ? (W + 15) / 16 : (W + 7) / 8, tilesY = lattice ? (localRows + 7) / 8 : (localRows + 3) / 4;
  const uint64_t total = (uint64_t)tilesX * tilesY * 32 * spp;
  std::vector<std::vector<uint8_t>> rows(tilesY);
  std::vector<std::vector<uint32_t>> lens(tilesY);
#pragma omp parallel for schedule(dynamic, 1)
  for (long ty = 0; ty < (long)tilesY; ++ty) {
    auto& out = rows[ty]; auto& ln = lens[ty];
    for (uint32_t tx = 0; tx < tilesX; ++tx) for (uint32_t k = 0; k < spp; ++k) for (uint32_t within = 0; within < 32; ++within) {
      const uint32_t x = lattice ? tx * 16 + 2 * (within & 7) + (k & 1) : tx * 8 + (within & 7), y = lattice ? ty * 8 + 2 * (within >> 3) + (k >> 1) : ty * 4 + (within >> 3);
      if (x >= W || y >= localRows) { ln.push_back(0); continue; }
      const uint32_t strip = y / stripRows;
      const uint32_t gy = (strip * stripStride + stripFirst) * stripRows + (y - strip * stripRows);
      const int si = lattice ? 0 : k / cam.nIter, sj = lattice ? 0 : k % cam.nIter;
      Slot s; Frame stack[RT_MAX_STACK]; Counters ctr; memset(&ctr, 0, sizeof ctr); memset(&s, 0, sizeof s);
      start_task(s, ctr, cam, x, gy, 0, si, sj);
      uint32_t q = 0;
      for (;;) {
        ++q;
        out.push_back((uint8_t)(s.kind | (s.ndirs << 2)));
        ShadowGeo sg;
        if (s.kind == K_TRACE) { s.minT = 1000.f; s.hitIdx = -1; if (vdot(s.rayD, s.rayD) != 0.f) for (uint32_t i = 0; i < sc.n; ++i) resolve_trace(s.minT, s.hitIdx, s.qo, s.rayD, sc.geo[i], i); }
        else if (s.kind == K_SHADOW) { s.blocked = 0; shadow_geo(s, sc, sg); for (int kk = 0; kk < s.ndirs; ++kk) for (uint32_t i = 0; i < sc.n; ++i) if (resolve_shadow(s.P, sg.d[kk], sg.gap[kk], sc.geo[i])) { s.blocked |= 1u << kk; break; } }
        else { s.hitIdx = -1; for (uint32_t i = 0; i < sc.n; ++i) resolve_contain(s.hitIdx, s.qo, sc.geo[i], i); }
        if (advance(s, stack, ctr, sc, cam, &sg)) break;
      }
      ln.push_back(q);
    }
  }
  FILE* f = fopen(path, "wb");
  Frame_ F = {n, W, H, spp, (uint32_t)S, localRows, tilesX, tilesY};
  fwrite(&F, sizeof F, 1, f);
  for (auto& ln : lens) fwrite(ln.data(), 4, ln.size(), f);
  for (auto& r : rows) fwrite(r.data(), 1, r.size(), f);
  fclose(f);
  uint64_t tq = 0; for (auto& r : rows) tq += r.size();
  printf("generated %llu work items, %llu queries\n", (unsigned long long)total, (unsigned long long)tq);
}

static void load(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(1); }
  if (fread(&gF, sizeof gF, 1, f) != 1) exit(1);
  const uint64_t total = (uint64_t)gF.tilesX * gF.tilesY * 32 * gF.spp;
  std::vector<uint32_t> len(total);
  if (fread(len.data(), 4, total, f) != total) exit(1);
  gOff.resize(total + 1); gOff[0] = 0;
  for (uint64_t i = 0; i < total; ++i) gOff[i + 1] = gOff[i] + len[i];
  gSeq.resize(gOff[total]);
  if (fread(gSeq.data(), 1, gSeq.size(), f) != gSeq.size()) exit(1);
  fclose(f);
}

/* ---------------- the model ---------------- */
struct Params {
  int nsm = 148, ctasPerSM = 2, warpsPerCta = 8, nslots = 4;
  double cycT = 19.0, cycS = 25.0, cycC = 11.0;   /* issue cycles per sphere of each pass                    */
  double fixW = 1500, fixL = 1500;                /* per-pass issue work / pure latency outside the loop     */
  double advW = 600, advL = 800;                  /* per advance round (two per trace/contain pass, one per shadow pass) */
  double ghz = 1.965;
  int order = 0;            /* 0 queue order | 1 non-empty tiles first | 2 tiles by decreasing total length | 3 by decreasing max length */
  int vote = 0;             /* 0 greedy fill | 1 oldest-first when the queue is dry | 2 fair: kind starving > starve passes gets served */
  int starve = 4;
  int split = 0;            /* 1: an under-filled pass costs loop * max(1/16, filled share) (split rounds)            */
  int mixed = 0;            /* 1: when dry, one shadow-form pass serves the first pending query of every lane regardless of kind */
  int pool = 0;             /* 1: cross-warp consolidation when dry (warps with < poolBelow live samples donate them all)   */
  int poolBelow = 64;
  int chunk = 0;
  int coarse = 8, classes = 16, deepAt = 32;
  int endSlots = 0;         /* >0: lanes refill only this many slots once fewer than endFrac of the work items remain */
  double endFrac = 0.02;
};

struct Slot_ { int64_t item = -1; uint32_t pos = 0; };
struct Warp {
  std::vector<Slot_> slots;   /* 32 * nslots, slot k of lane l at l * nslots + k */
  uint32_t wbase = 0, wend = 0;
  bool dry = false, done = false, phase2 = false, starved = false; uint32_t tile = 0;
  double dryAt = 0, endAt = 0;
  int sched = 0;
  int starveS = 0, starveC = 0, starveT = 0;
  uint64_t passes = 0, served = 0, cap = 0;
};

int main(int argc, char** argv) {
  if (argc >= 11 && !strcmp(argv[1], "gen")) { gen(argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atof(argv[6]), atoi(argv[7]), atoi(argv[8]), atoi(argv[9]), atoi(argv[10]), argc > 11 ? atoi(argv[11]) : 0); return 0; }
  if (argc < 3) { fprintf(stderr, "usage\n"); return 1; }
  load(argv[2]);
  Params P;
  std::map<std::string, double*> dk = {{"cycT", &P.cycT}, {"cycS", &P.cycS}, {"cycC", &P.cycC}, {"fixW", &P.fixW}, {"fixL", &P.fixL}, {"advW", &P.advW}, {"advL", &P.advL}, {"endFrac", &P.endFrac}};
  std::map<std::string, int*> ik = {{"nslots", &P.nslots}, {"order", &P.order}, {"vote", &P.vote}, {"starve", &P.starve}, {"split", &P.split}, {"mixed", &P.mixed}, {"pool", &P.pool},
                                    {"poolBelow", &P.poolBelow}, {"chunk", &P.chunk}, {"endSlots", &P.endSlots}, {"coarse", &P.coarse}, {"deepAt", &P.deepAt}, {"classes", &P.classes}, {"ctasPerSM", &P.ctasPerSM}, {"nsm", &P.nsm}};
  for (int i = 3; i < argc; ++i) {
    char* eq = strchr(argv[i], '='); if (!eq) continue;
    std::string k(argv[i], eq - argv[i]);
    if (dk.count(k)) *dk[k] = atof(eq + 1); else if (ik.count(k)) *ik[k] = atoi(eq + 1); else { fprintf(stderr, "unknown key %s\n", k.c_str()); return 1; }
  }
  const uint64_t totalWork = (uint64_t)gF.tilesX * gF.tilesY * 32 * gF.spp;
  const int nWarps = P.nsm * P.ctasPerSM * P.warpsPerCta, nSched = P.nsm * 4;
  uint32_t chunk = P.chunk ? P.chunk : (uint32_t)((totalWork / ((uint64_t)nWarps * 16)) & ~31ull);
  if (!P.chunk) { if (chunk < 32) chunk = 32; if (chunk > 256) chunk = 256; }
  /* work order: permutation of 32-item groups (tile, sample) */
  const uint64_t nGroups = totalWork / 32;
  std::vector<uint32_t> perm(nGroups);
  for (uint64_t g = 0; g < nGroups; ++g) perm[g] = (uint32_t)g;
  if (P.order) {
    const uint64_t nTiles = nGroups / gF.spp;
    std::vector<double> key(nTiles);
    for (uint64_t t = 0; t < nTiles; ++t) {
      uint64_t tot = 0, mx = 0;
      for (uint64_t i = t * gF.spp * 32; i < (t + 1) * gF.spp * 32; ++i) { const uint64_t l = gOff[i + 1] - gOff[i]; tot += l; mx = std::max(mx, l); }
      /* the probe sample: pixel (3,1) of the tile, sample 0 */
      const uint64_t pi = t * gF.spp * 32 + 11;
      const uint64_t pl = gOff[pi + 1] - gOff[pi];
      bool refr = false; for (uint64_t q = gOff[pi]; q < gOff[pi + 1]; ++q) if ((gSeq[q] & 3) == 3) refr = true;
      key[t] = P.order == 1 ? (tot > gF.spp * 32 ? 1.0 : 0.0) : P.order == 2 ? (double)tot : P.order == 3 ? (double)mx
             : P.order == 4 ? (pl <= 1 ? 0.0 : refr ? 2.0 : 1.0) : P.order == 5 ? (double)pl : P.order == 6 ? (double)std::min<uint64_t>(pl, 40) / 8 : 0.0;
    }
    std::vector<uint32_t> tiles(nTiles);
    for (uint64_t t = 0; t < nTiles; ++t) tiles[t] = (uint32_t)t;
    if (P.order == 7) {
      /* coarse-then-LPT: tile rows = 0 mod R first (queue order), the rest by the maximum chain length seen in the
       * bracketing coarse rows (tx-1..tx+1), quantised to `classes` buckets, deepest first */
      const int R = P.coarse;
      std::vector<double> mxT(nTiles);
      for (uint64_t t = 0; t < nTiles; ++t) { uint64_t mx = 0; for (uint64_t i = t * gF.spp * 32; i < (t + 1) * gF.spp * 32; ++i) mx = std::max<uint64_t>(mx, gOff[i + 1] - gOff[i]); mxT[t] = (double)mx; }
      for (uint64_t t = 0; t < nTiles; ++t) {
        const int ty = (int)(t / gF.tilesX), tx = (int)(t % gF.tilesX);
        if (ty % R == 0) { key[t] = 1e9 - (double)t * 1e-3; continue; }
        const int y0 = ty / R * R, y1 = std::min<int>(y0 + R, (int)gF.tilesY - 1) / R * R;
        double m = 0;
        for (int yy : {y0, y1}) for (int xx = std::max(0, tx - 1); xx <= std::min<int>(gF.tilesX - 1, tx + 1); ++xx) m = std::max(m, mxT[(uint64_t)yy * gF.tilesX + xx]);
        key[t] = floor(std::min(m, 127.0) / (128.0 / P.classes));
      }
    }
    std::stable_sort(tiles.begin(), tiles.end(), [&](uint32_t a, uint32_t b) { return key[a] > key[b]; });
    for (uint64_t t = 0; t < nTiles; ++t) for (uint32_t k = 0; k < gF.spp; ++k) perm[t * gF.spp + k] = tiles[t] * gF.spp + k;
    if (P.order == 8) {
      /* sample 0 of every tile first (queue order), then the other samples of the tiles by the maximum chain length
       * their sample 0 showed, deepest first (quantised to `classes` buckets) */
      std::vector<double> k0(nTiles);
      for (uint64_t t = 0; t < nTiles; ++t) { uint64_t mx = 0; for (uint64_t i = t * gF.spp * 32; i < t * gF.spp * 32 + 32; ++i) mx = std::max<uint64_t>(mx, gOff[i + 1] - gOff[i]); k0[t] = floor(std::min<double>(mx, 127.0) / (128.0 / P.classes)); }
      for (uint64_t t = 0; t < nTiles; ++t) tiles[t] = (uint32_t)t;
      std::stable_sort(tiles.begin(), tiles.end(), [&](uint32_t a, uint32_t b) { return k0[a] > k0[b]; });
      uint64_t o = 0;
      for (uint64_t t = 0; t < nTiles; ++t) perm[o++] = (uint32_t)(t * gF.spp);
      for (uint64_t t = 0; t < nTiles; ++t) for (uint32_t k = 1; k < gF.spp; ++k) perm[o++] = tiles[t] * gF.spp + k;
    }
  }
  const double nPad = (gF.n + 31) / 32 * 32;
  std::vector<Warp> warps(nWarps);
  for (int w = 0; w < nWarps; ++w) {
    warps[w].slots.resize(32 * P.nslots);
    const int cta = w / P.warpsPerCta, sm = cta % P.nsm;
    warps[w].sched = sm * 4 + (w % P.warpsPerCta) % 4;
  }
  uint64_t qhead = 0;
  /* order=9: dynamic buckets.  Phase 1 = sample 0 of every tile from the queue; a tile's other samples are pushed into a
   * bucket when one of its phase-1 samples reaches `deepAt` queries (bucket 0) or when all 32 have completed (by class) */
  const bool dyn = P.order == 9 || P.order == 10;
  const bool sweep = P.order == 10;     /* only the deep trigger pushes; everything else is swept in tile order */
  uint64_t sweepNext = 0;
  const uint64_t nTilesD = nGroups / gF.spp;
  std::vector<uint8_t> tPushed(dyn ? nTilesD : 0, 0), tDone(dyn ? nTilesD : 0, 0), tMax(dyn ? nTilesD : 0, 0);
  std::vector<std::vector<uint32_t>> bucket(8); std::vector<size_t> bhead(8, 0);
  uint64_t tilesConsumed = 0;
  auto class_of = [&](unsigned mx) { return mx >= (unsigned)P.deepAt ? 0 : mx >= 16 ? 1 : mx >= 8 ? 2 : mx >= 2 ? 3 : 4; };
  std::vector<std::pair<int64_t, uint32_t>> pool;   /* donated samples (item, pos) */
  /* processor sharing per scheduler: each warp alternates [issue work W (shared)] -> [latency L (private)] */
  struct Job { double W, L; };
  std::vector<double> remW(nWarps, 0.0), lat(nWarps, 0.0);
  std::vector<int> state(nWarps, 0);               /* 0 needs a new pass, 1 in issue work, 2 in latency (wake time in lat) */
  std::vector<std::vector<int>> onSched(nSched);
  for (int w = 0; w < nWarps; ++w) onSched[warps[w].sched].push_back(w);
  std::vector<double> schedClock(nSched, 0.0);
  uint64_t totPasses = 0, totServed = 0, totCap = 0, passK[4] = {0, 0, 0, 0};
  double busyW = 0;

  auto live_count = [&](Warp& w) { int c = 0; for (auto& s : w.slots) c += s.item >= 0; return c; };
  /* decide and account the next pass of warp w at time `now`; returns false when the warp is finished */
  auto next_pass = [&](int wi, double now, Job& job) -> bool {
    Warp& w = warps[wi];
    /* refill */
    const bool nearEnd = P.endSlots > 0 && (double)(totalWork - std::min(totalWork, qhead)) < P.endFrac * (double)totalWork;
    w.starved = false;
    for (int k = 0; k < P.nslots && !w.dry && !w.starved; ++k) {
      if (nearEnd && k >= P.endSlots) break;
      for (int l = 0; l < 32 && !w.dry && !w.starved; ++l) {
        Slot_& s = w.slots[l * P.nslots + k];
        while (s.item < 0 && !w.starved) {
          if (P.pool && qhead >= totalWork && !pool.empty()) { s.item = pool.back().first; s.pos = pool.back().second; pool.pop_back(); break; }
          if (w.wbase >= w.wend && dyn) {
            if (qhead < nTilesD * 32) { w.wbase = (uint32_t)qhead; w.wend = (uint32_t)std::min<uint64_t>(qhead + chunk, nTilesD * 32); qhead += chunk; w.phase2 = false; }
            else {
              int b = 0; while (b < 8 && bhead[b] >= bucket[b].size()) ++b;
              uint32_t t;
              if (b == 8 && sweep) {
                while (sweepNext < nTilesD && tPushed[sweepNext]) ++sweepNext;
                if (sweepNext >= nTilesD) { w.dry = true; w.dryAt = now; w.starved = true; break; }
                t = (uint32_t)sweepNext; tPushed[t] = 1; ++tilesConsumed;
              } else {
              if (b == 8) { if (tilesConsumed == nTilesD) { w.dry = true; w.dryAt = now; } w.starved = true; break; }
              t = bucket[b][bhead[b]++]; ++tilesConsumed;
              }
              w.wbase = 0; w.wend = 32 * (gF.spp - 1); w.phase2 = true; w.tile = t;
            }
          } else
          if (w.wbase >= w.wend) {
            if (qhead >= totalWork) { w.dry = true; w.dryAt = now; break; }
            w.wbase = (uint32_t)qhead; w.wend = (uint32_t)std::min<uint64_t>(qhead + chunk, totalWork); qhead += chunk;
          }
          const uint64_t idx = w.wbase++;
          const uint64_t item = dyn ? (w.phase2 ? ((uint64_t)w.tile * gF.spp + 1 + (idx >> 5)) * 32 + (idx & 31) : (uint64_t)(idx >> 5) * gF.spp * 32 + (idx & 31))
                                    : (uint64_t)perm[idx >> 5] * 32 + (idx & 31);
          if (gOff[item + 1] > gOff[item]) { s.item = (int64_t)item; s.pos = 0; }
        }
      }
    }
    if (w.dry && P.pool) {
      /* take donated samples into free slots; donate everything when under-filled and somebody else can take it */
      for (auto& s : w.slots) if (s.item < 0 && !pool.empty()) { s.item = pool.back().first; s.pos = pool.back().second; pool.pop_back(); }
    }
    /* census */
    int nT = 0, nS = 0, nC = 0, ndMax = 0;
    std::vector<int> t0(32, -1), t1(32, -1), s0(32, -1), c0(32, -1), c1(32, -1);
    for (int l = 0; l < 32; ++l) for (int k = 0; k < P.nslots; ++k) {
      const Slot_& s = w.slots[l * P.nslots + k];
      if (s.item < 0) continue;
      const uint8_t q = gSeq[gOff[s.item] + s.pos]; const int kind = q & 3, nd = q >> 2;
      if (kind == 1) { if (t0[l] < 0) t0[l] = k; else if (t1[l] < 0) t1[l] = k; }
      else if (kind == 2) { if (s0[l] < 0) { s0[l] = k; ndMax = std::max(ndMax, nd); } }
      else { if (c0[l] < 0) c0[l] = k; else if (c1[l] < 0) c1[l] = k; }
    }
    int lanesS = 0;
    for (int l = 0; l < 32; ++l) {
      nT += (t0[l] >= 0) + (t1[l] >= 0); nC += (c0[l] >= 0) + (c1[l] >= 0);
      if (s0[l] >= 0) { nS += gSeq[gOff[w.slots[l * P.nslots + s0[l]].item] + w.slots[l * P.nslots + s0[l]].pos] >> 2; ++lanesS; }
      else if (t0[l] >= 0) nS += 1;
    }
    if (nT + nC + lanesS == 0) {
      if (w.dry && (!P.pool || pool.empty())) { w.done = true; w.endAt = now; return false; }
      if (dyn) { job.W = 50; job.L = 2000; return true; }
      job.W = 200; job.L = 0; return true;     /* spin once */
    }
    if (w.dry && P.pool && live_count(w) < P.poolBelow) {
      /* donate all and retry as a consumer next round (the donor itself may take them back) */
      bool othersAlive = false;
      for (int o = 0; o < nWarps && !othersAlive; ++o) if (o != wi && !warps[o].done) othersAlive = true;
      if (othersAlive && (int)pool.size() + live_count(w) >= 0) {
        int lc = live_count(w);
        /* only donate if the pool plus ours could fill somebody better: simple rule — donate when pool non-empty or ours < poolBelow/2 */
        (void)lc;
      }
    }
    const unsigned capS = (ndMax <= 2) ? 64u : 128u;
    const bool anyS = ndMax > 0;
    int mode = (anyS && (unsigned)nS * 64u >= (unsigned)nT * capS && (unsigned)nS * 64u >= (unsigned)nC * capS) ? 2 : (nT >= nC) ? 1 : 3;
    if (P.vote == 2) {
      if (lanesS && w.starveS >= P.starve) mode = 2; else if (nC && w.starveC >= P.starve) mode = 3; else if (nT && w.starveT >= P.starve) mode = 1;
    }
    const bool mixed = P.mixed && w.dry;
    int served = 0, capacity = 0, advRounds = 0;
    double loopCyc = 0;
    auto step = [&](int l, int k) {
      Slot_& s = w.slots[l * P.nslots + k]; ++s.pos; ++served;
      const bool fin = s.pos >= gOff[s.item + 1] - gOff[s.item];
      if (dyn && ((s.item >> 5) % gF.spp) == 0) {
        const uint64_t t = (s.item >> 5) / gF.spp;
        tMax[t] = (uint8_t)std::max<unsigned>(tMax[t], std::min<unsigned>(s.pos, 255));
        if (fin) tDone[t]++;
        if (!tPushed[t] && ((int)s.pos >= P.deepAt || (!sweep && tDone[t] == 32))) { tPushed[t] = 1; bucket[sweep ? 0 : class_of(tMax[t])].push_back((uint32_t)t); }
      }
      if (fin) s.item = -1;
    };
    if (mixed) {
      for (int l = 0; l < 32; ++l) { const int k = s0[l] >= 0 ? s0[l] : t0[l] >= 0 ? t0[l] : c0[l]; if (k >= 0) step(l, k); }
      capacity = 32; loopCyc = P.cycS * nPad; advRounds = 1; passK[2]++;
    } else if (mode == 2) {
      for (int l = 0; l < 32; ++l) { const int k = s0[l] >= 0 ? s0[l] : t0[l]; if (k >= 0) step(l, k); }
      capacity = 32; loopCyc = P.cycS * nPad; advRounds = 1; passK[2]++;
      w.starveS = 0; if (nC) w.starveC++; if (nT) w.starveT++;
    } else if (mode == 1) {
      for (int l = 0; l < 32; ++l) { if (t0[l] >= 0) step(l, t0[l]); if (t1[l] >= 0) step(l, t1[l]); }
      capacity = 64; loopCyc = P.cycT * nPad; advRounds = 2; passK[1]++;
      w.starveT = 0; if (nC) w.starveC++; if (lanesS) w.starveS++;
    } else {
      for (int l = 0; l < 32; ++l) { if (c0[l] >= 0) step(l, c0[l]); if (c1[l] >= 0) step(l, c1[l]); }
      capacity = 64; loopCyc = P.cycC * nPad; advRounds = 2; passK[3]++;
      w.starveC = 0; if (nT) w.starveT++; if (lanesS) w.starveS++;
    }
    if (P.split) { const double share = std::max(1.0 / 16, (double)served / capacity); double f = 1; while (f * 2 * share <= 1.0 && f < 16) f *= 2; loopCyc = loopCyc / f + 40 * log2(f); }
    if (advRounds == 2 && served <= 32) { bool two = false; for (int l = 0; l < 32; ++l) if ((mode == 1 ? t1[l] : c1[l]) >= 0) two = true; if (!two) advRounds = 1; }
    job.W = P.fixW + loopCyc + advRounds * P.advW;
    job.L = P.fixL + advRounds * P.advL;
    w.passes++; w.served += served; w.cap += capacity; totPasses++; totServed += served; totCap += capacity;
    return true;
  };

  /* event loop: per scheduler processor sharing.  Global time advances scheduler by scheduler via a priority queue of next events */
  typedef std::pair<double, int> Ev;   /* (time, scheduler) */
  std::priority_queue<Ev, std::vector<Ev>, std::greater<Ev>> pq;
  for (int s = 0; s < nSched; ++s) pq.push({0.0, s});
  double tEnd = 0;
  while (!pq.empty()) {
    const Ev ev = pq.top(); pq.pop();
    const int s = ev.second; const double now = ev.first;
    /* advance the issue work of this scheduler's running warps from schedClock[s] to now */
    int running = 0;
    for (int wi : onSched[s]) if (state[wi] == 1) ++running;
    if (running) { const double dt = (now - schedClock[s]) / running; for (int wi : onSched[s]) if (state[wi] == 1) remW[wi] -= dt; busyW += now - schedClock[s]; }
    schedClock[s] = now;
    for (int wi : onSched[s]) {
      if (warps[wi].done) continue;
      if (state[wi] == 1 && remW[wi] <= 1e-6) { state[wi] = 2; lat[wi] = now + lat[wi]; }
      if (state[wi] == 2 && lat[wi] <= now + 1e-9) state[wi] = 0;
      if (state[wi] == 0) {
        Job j;
        if (next_pass(wi, now, j)) { state[wi] = 1; remW[wi] = j.W; lat[wi] = j.L; }
        else tEnd = std::max(tEnd, now);
      }
    }
    /* next event of this scheduler */
    running = 0; double minW = 1e300, minL = 1e300;
    for (int wi : onSched[s]) { if (warps[wi].done) continue; if (state[wi] == 1) { ++running; minW = std::min(minW, remW[wi]); } else if (state[wi] == 2) minL = std::min(minL, lat[wi]); }
    double next = 1e300;
    if (running) next = now + std::max(minW, 0.0) * running;
    next = std::min(next, minL);
    if (next < 1e299) pq.push({std::max(next, now + 1e-3), s});
  }
  double meanTail = 0, maxTail = 0, meanDry = 0;
  for (auto& w : warps) { meanTail += w.endAt - w.dryAt; maxTail = std::max(maxTail, w.endAt - w.dryAt); meanDry += w.dryAt; }
  meanTail /= nWarps; meanDry /= nWarps;
  const double ms = tEnd / (P.ghz * 1e6);
  printf("kernel_ms %.3f  dry_at_ms %.3f  mean_tail_ms %.3f  max_tail_ms %.3f  passes %llu (T %llu S %llu C %llu)  fill %.4f  issue_busy %.3f  queries %llu\n", ms, meanDry / (P.ghz * 1e6),
         meanTail / (P.ghz * 1e6), maxTail / (P.ghz * 1e6), (unsigned long long)totPasses, (unsigned long long)passK[1], (unsigned long long)passK[2], (unsigned long long)passK[3],
         (double)totServed / totCap, busyW / (tEnd * nSched), (unsigned long long)gSeq.size());
  return 0;
}
