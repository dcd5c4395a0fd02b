"""GPU parity tests (run on a B200 through gpurun): the CUDA path, called through the
C-ABI, against the oracle on the same inputs and against the committed golden vectors.

Bar: the float framebuffer is BIT-EXACT (NaN-ness equal, every other value the same
bits), which implies north_star's tolerance (<= 1 LSB per 8-bit channel on >= 99.9 % of
pixels) with zero offenders; the 8-bit comparison is asserted explicitly as well.
Nothing here reads /root/reference.
"""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).resolve().parent / "golden"
FACTS = json.loads((GOLD / "facts.json").read_text())
G = np.load(GOLD / "golden.npz")
FB_CASES = [k for k in FACTS if "W" in FACTS[k]]
TOL_LSB = 1          # north_star tolerance, per 8-bit channel
TOL_FRAC = 0.999     # on at least this fraction of pixels


def _case_scene(pkg, case):
    name = case.split("_")[0]
    sph = np.ascontiguousarray(G[f"{name}_spheres"]).view(pkg.SPHERE_DTYPE).reshape(-1)
    lgt = np.ascontiguousarray(G[f"{name}_lights"]).view(pkg.LIGHT_DTYPE).reshape(-1)
    return sph, lgt


@pytest.fixture(scope="module")
def gpu(pkg):
    assert pkg.device_count() > 0, "no CUDA device: the GPU tests need a B200"
    r = pkg.Renderer(0)
    yield r
    r.close()


def _render(gpu, sph, lgt, W, H, zoom, alias, S, **opts):
    for k, v in opts.items():
        gpu.set_option(k, v)
    gpu.upload_scene(sph, lgt)
    gpu.render(W, H, zoom, alias, S)
    fb, mx = gpu.readback()
    st = gpu.stats()
    for k in opts:
        gpu.set_option(k, 0)
    return fb, mx, st


def _assert_parity(orc_mod, oracle, ref, got):
    rep = orc_mod.compare(ref, got, oracle)
    assert rep["nan_masks_equal"], rep
    assert rep["within_1lsb_frac"] >= TOL_FRAC and rep["max_lsb_diff"] <= TOL_LSB, rep
    assert rep["bit_exact"], rep
    return rep


@pytest.mark.parametrize("case", FB_CASES)
def test_golden_framebuffers(pkg, orc_mod, gpu, case):
    f = FACTS[case]
    sph, lgt = _case_scene(pkg, case)
    fb, _, _ = _render(gpu, sph, lgt, f["W"], f["H"], f["zoom"], f["alias"], f["S"])
    assert np.array_equal(orc_mod.canon(fb), G[case])


@pytest.mark.parametrize("S", [5, 6])
def test_config1_default_scene_full_frame(pkg, orc_mod, oracle, gpu, S):
    """BASELINE config 1: the reference default scene at its default resolution (800x600,
    alias 3), stack 6 (CPU copy) and 5 (OpenCL copy's stack size)."""
    f = FACTS[f"default_800x600_a3_s{S}"]
    sph, lgt = pkg.default_scene()
    fb, mx, st = _render(gpu, sph, lgt, 800, 600, -4.0, 3.0, S)
    assert hashlib.md5(orc_mod.canon(fb).tobytes()).hexdigest() == f["float_md5_canon"]
    assert int(np.float32(mx).view(np.uint32)) == f["max_bits"]
    assert int(np.isnan(fb).any(axis=2).sum()) == f["nan_pixels"]
    # the device quantiser reproduces the reference PPM byte for byte
    rgb = gpu.readback_rgb8()
    ppm = b"P6\n800 600\n255\n" + rgb.tobytes()
    assert hashlib.md5(ppm).hexdigest() == f["ppm_md5"]
    ref, ctr = oracle.render(sph, lgt, 800, 600, -4.0, 3.0, S)
    _assert_parity(orc_mod, oracle, ref, fb)
    for k in ("rays", "shadow_rays", "contain_queries", "contain_tests", "samples"):
        assert st[k] == ctr[k], k
    assert st["kernel_launches"] >= 1


def test_config2_default_scene_1080p(pkg, orc_mod, oracle, gpu):
    """BASELINE config 2: default scene, 1920x1080, 1 spp, depth 4."""
    sph, lgt = pkg.default_scene()
    fb, mx, st = _render(gpu, sph, lgt, 1920, 1080, -4.0, 1.0, 4)
    ref, ctr = oracle.render(sph, lgt, 1920, 1080, -4.0, 1.0, 4)
    _assert_parity(orc_mod, oracle, ref, fb)
    assert st["rays"] == ctr["rays"] and mx == oracle.max_colour(ref)


@pytest.mark.parametrize("n,l,seed,W,H,alias,S", [
    (256, 4, 0, 480, 270, 1.0, 6),      # config 3 scene at 1/8 linear size
    (1024, 4, 0, 240, 135, 2.0, 8),     # config 4 scene at 1/32 linear size
    (16, 4, 0, 320, 180, 1.0, 8),       # config 5 sweep ends
    (4096, 4, 0, 96, 54, 1.0, 8),
    (100, 3, 11, 211, 97, 1.5, 7),      # ragged: odd sizes, fractional alias, 3 lights
])
def test_synthetic_scenes_match_oracle(pkg, orc_mod, oracle, gpu, n, l, seed, W, H, alias, S):
    sph, lgt = pkg.synth_scene(n, l, seed=seed)
    fb, mx, st = _render(gpu, sph, lgt, W, H, -4.0, alias, S)
    ref, ctr = oracle.render(sph, lgt, W, H, -4.0, alias, S)
    _assert_parity(orc_mod, oracle, ref, fb)
    assert st["rays"] == ctr["rays"] and st["contain_tests"] == ctr["contain_tests"]
    assert mx == oracle.max_colour(ref)
    rgb = gpu.readback_rgb8()
    assert np.array_equal(rgb, oracle.quantise(ref, oracle.max_colour(ref)))


def test_staging_and_filter_variants_agree(pkg, orc_mod, gpu):
    """__constant__ vs shared-memory (TMA bulk) staging and the filter switch are
    invisible in the output."""
    sph, lgt = pkg.synth_scene(200, 4, seed=3)
    base, _, st0 = _render(gpu, sph, lgt, 160, 90, -4.0, 1.0, 8)
    for opts in ({"staging": 1}, {"staging": 2}, {"no_filter": 1}, {"blocks_per_sm": 1}, {"staging": 1, "no_filter": 1},
                 {"accel": 2}, {"accel": 2, "no_filter": 1}):
        fb, _, st = _render(gpu, sph, lgt, 160, 90, -4.0, 1.0, 8, **opts)
        assert np.array_equal(orc_mod.canon(fb), orc_mod.canon(base)), opts
        assert st["rays"] == st0["rays"]
    _, _, st1 = _render(gpu, sph, lgt, 160, 90, -4.0, 1.0, 8, staging=1)
    _, _, st2 = _render(gpu, sph, lgt, 160, 90, -4.0, 1.0, 8, staging=2)
    assert st1["staging"] == 1 and st2["staging"] == 2


def test_work_order_and_lockstep_variants_agree(pkg, orc_mod, oracle, gpu):
    """The work order (deep tiles first / scanline), its tuning knobs, lockstep passes and the slot placement are
    scheduling only: same frame, same ray count, for 1 spp (pixel sub-lattices) and for several samples per pixel,
    on a frame large enough for every queue phase (first groups, deep list, sweep) to be exercised."""
    sph, lgt = pkg.synth_scene(300, 4, seed=5)
    for W, H, alias in ((601, 403, 1.0), (320, 200, 2.0), (97, 61, 3.0)):
        ref, ctr = oracle.render(sph, lgt, W, H, -4.0, alias, 8)
        for opts in ({}, {"order": 1}, {"order": 2}, {"order": 1, "lockstep": 1}, {"order": 1, "lockstep": 2}, {"order": 1, "lockstep": 3}, {"lockstep": 3}, {"order": 2, "lockstep": 1},
                     {"order": 1, "sweep_step": 3, "deep_at": 6}, {"order": 1, "deep_at": 2, "sweep_step": 8, "lockstep": 2},
                     {"order": 1, "slot_mode": 2, "deep_at": 4}, {"order": 1, "accel": 2, "deep_at": 3}):
            fb, _, st = _render(gpu, sph, lgt, W, H, -4.0, alias, 8, **opts)
            assert np.array_equal(orc_mod.canon(fb), orc_mod.canon(ref)), (W, H, alias, opts)
            assert st["rays"] == ctr["rays"] and st["samples"] == ctr["samples"], (W, H, alias, opts)


def test_work_order_protocol_stress(pkg, orc_mod, gpu):
    """The deep list and the sweep hand every tile out exactly once whatever the timing: random frame shapes (ragged
    tiles, strips, 1 and 4 spp) with the deep trigger at 1-3 queries (every tile is listed by its first samples), claims
    of 1-8 tiles and every lockstep mode, against the scanline order of the same frame: same bits, same ray and sample
    counts."""
    rng = np.random.default_rng(7)
    sph, lgt = pkg.synth_scene(96, 3, seed=9)
    for it in range(24):
        W, H = int(rng.integers(9, 400)), int(rng.integers(5, 300))
        alias = float(rng.choice([1.0, 2.0]))
        strips = (int(rng.integers(1, 9)), 0, 1) if it % 3 else (4, int(rng.integers(0, 3)), 3)
        opts = {"order": 1, "deep_at": int(rng.integers(1, 4)), "sweep_step": int(rng.integers(1, 9)), "lockstep": int(rng.integers(0, 4))}
        frames = []
        for o in ({"order": 2}, opts):
            for k, v in o.items():
                gpu.set_option(k, v)
            gpu.upload_scene(sph, lgt)
            if it % 3:
                gpu.render(W, H, -4.0, alias, 8)
            else:
                gpu.render_strips(W, H, -4.0, alias, 8, *strips)
            fb, _ = gpu.readback()
            st = gpu.stats()
            for k in o:
                gpu.set_option(k, 0)
            frames.append((orc_mod.canon(fb), st["rays"], st["samples"]))
        assert np.array_equal(frames[0][0], frames[1][0]), (it, W, H, alias, strips, opts)
        assert frames[0][1:] == frames[1][1:], (it, W, H, alias, strips, opts)


def test_edge_cases(pkg, orc_mod, oracle, gpu):
    sph, lgt = pkg.default_scene()
    # empty scene, no lights, 1-pixel-high and non-tile-multiple frames, sub-unit alias
    fb, mx, st = _render(gpu, sph[:0], lgt, 37, 5, -4.0, 1.0, 6)
    assert not fb.any() and mx == 1.0 and st["rays"] == 37 * 5
    for W, H, alias, S, s, l in [(33, 1, 1.0, 6, sph, lgt), (8, 4, 3.0, 1, sph, lgt), (61, 47, 0.5, 6, sph, lgt),
                                 (64, 48, 2.0, 6, sph, lgt[:0]), (50, 50, 1.0, 16, sph, lgt)]:
        fb, _, _ = _render(gpu, s, l, W, H, -4.0, alias, S)
        ref, _ = oracle.render(s, l, W, H, -4.0, alias, S)
        _assert_parity(orc_mod, oracle, ref, fb)
    # argument validation through the C-ABI
    lib = pkg.load()
    assert lib.rt_cuda_render(gpu._ctx, 0, 10, -4.0, 1.0, 6) == -1
    assert lib.rt_cuda_render(gpu._ctx, 10, 10, -4.0, 1.0, 0) == -1
    assert lib.rt_cuda_render(gpu._ctx, 10, 10, -4.0, 1.0, 17) == -1
    big = np.zeros(20000, pkg.SPHERE_DTYPE)
    assert lib.rt_cuda_upload_scene(gpu._ctx, big.ctypes.data, len(big), None, 0) == -6
    fresh = pkg.Renderer(0)
    assert lib.rt_cuda_render(fresh._ctx, 8, 8, -4.0, 1.0, 6) == -4          # no scene yet
    fresh.upload_scene(sph, lgt)
    out = np.zeros((8, 8, 3), np.float32)
    assert lib.rt_cuda_readback(fresh._ctx, out.ctypes.data, None) == -5     # no frame yet
    fresh.close()


def test_accelerated_mode_matches_oracle(pkg, orc_mod, oracle, gpu):
    """Option accel=1 (two-level cluster filter) changes which spheres are looked
    at, never the answer: bit-exact frames and the reference's work counters, far fewer filter tests."""
    for (sph, lgt), (W, H, alias, S) in [(pkg.synth_scene(300, 4, seed=4), (160, 90, 2.0, 8)),
                                          (pkg.synth_scene(1024, 4), (192, 108, 1.0, 8)),
                                          (pkg.synth_scene(64, 2, seed=8), (97, 61, 1.0, 5)),
                                          (pkg.synth_scene(2500, 3, seed=2), (96, 54, 1.0, 6))]:
        ref, ctr = oracle.render(sph, lgt, W, H, -4.0, alias, S)
        fb0, _, st0 = _render(gpu, sph, lgt, W, H, -4.0, alias, S)
        fb, mx, st = _render(gpu, sph, lgt, W, H, -4.0, alias, S, accel=2)
        assert st["accel"] == 1 and st0["accel"] == 0 and st["clusters"] >= (len(sph) + 7) // 8
        _assert_parity(orc_mod, oracle, ref, fb)
        assert np.array_equal(orc_mod.canon(fb), orc_mod.canon(fb0))
        for k in ("rays", "shadow_rays", "contain_queries", "contain_tests", "samples"):
            assert st[k] == ctr[k], k
        assert st["filter_tests"] * 2 < st0["filter_tests"]
    # too few spheres to cull: the option is ignored
    sph, lgt = pkg.default_scene()
    _, _, st = _render(gpu, sph, lgt, 64, 48, -4.0, 1.0, 6, accel=2)
    assert st["accel"] == 0
    # accel=1 only engages where it pays (>= 768 spheres)
    sph, lgt = pkg.synth_scene(300, 4, seed=4)
    assert _render(gpu, sph, lgt, 64, 48, -4.0, 1.0, 6, accel=1)[2]["accel"] == 0
    sph, lgt = pkg.synth_scene(800, 4, seed=4)
    assert _render(gpu, sph, lgt, 64, 48, -4.0, 1.0, 6, accel=1)[2]["accel"] == 1


def test_strips_and_multisample_match_oracle(pkg, orc_mod, oracle, gpu):
    """Single- and multi-sample frames, whole frame and strips, against the oracle with its work counters."""
    for (sph, lgt), (W, H, alias, S) in [(pkg.default_scene(), (200, 150, 3.0, 6)),
                                          (pkg.synth_scene(300, 4, seed=4), (160, 90, 2.0, 8)),
                                          (pkg.synth_scene(64, 2, seed=8), (97, 61, 1.0, 5))]:
        fb, mx, st = _render(gpu, sph, lgt, W, H, -4.0, alias, S)
        assert st["engine"] == 1
        ref, ctr = oracle.render(sph, lgt, W, H, -4.0, alias, S)
        _assert_parity(orc_mod, oracle, ref, fb)
        for k in ("rays", "shadow_rays", "contain_queries", "contain_tests", "samples"):
            assert st[k] == ctr[k], k
        assert mx == oracle.max_colour(ref)
        gpu.render_strips(W, H, -4.0, alias, S, 8, 1, 3)
        part, _ = gpu.readback()
        assert np.array_equal(orc_mod.canon(part), orc_mod.canon(fb[pkg.local_rows(H, 8, 1, 3)]))


def test_two_contexts_interleaved(pkg, orc_mod, oracle):
    """Two live contexts on device 0 holding different small scenes, rendered in turn with every staging:
    each frame is its own scene's (the filter records of the constant-bank path travel with the launch,
    no module-global state is shared between contexts)."""
    a, b = pkg.Renderer(0), pkg.Renderer(0)
    try:
        sa, la = pkg.default_scene()
        sb, lb = pkg.synth_scene(40, 4, seed=2)
        ref_a, _ = oracle.render(sa, la, 160, 120, -4.0, 2.0, 6)
        ref_b, _ = oracle.render(sb, lb, 160, 120, -4.0, 1.0, 8)
        a.upload_scene(sa, la)
        b.upload_scene(sb, lb)            # uploaded AFTER a's scene, before a renders
        for staging in (0, 1, 2):
            a.set_option("staging", staging)
            b.set_option("staging", staging)
            for _ in range(2):
                a.render(160, 120, -4.0, 2.0, 6)
                b.render(160, 120, -4.0, 1.0, 8)       # both in flight on their own streams
                fa, _ = a.readback()
                fb, _ = b.readback()
                _assert_parity(orc_mod, oracle, ref_a, fa)
                _assert_parity(orc_mod, oracle, ref_b, fb)
            if staging:
                assert a.stats()["staging"] == staging and b.stats()["staging"] == staging
    finally:
        a.close()
        b.close()


def test_failed_render_leaves_no_frame(pkg, gpu):
    """A render that fails validation must not publish a frame: later readbacks say NO_FRAME
    or keep serving the previous, complete frame — never uninitialised pixels at the new size."""
    lib = pkg.load()
    sph, lgt = pkg.default_scene()
    fresh = pkg.Renderer(0)
    try:
        fresh.upload_scene(sph, lgt)
        assert lib.rt_cuda_render(fresh._ctx, 16, 16, -4.0, 65536.0, 6) == -6        # alias^2 samples do not fit 32 bits
        assert lib.rt_cuda_render(fresh._ctx, 16, 16, -4.0, 2.0e6, 6) == -6          # beyond the iteration cap
        out = np.zeros((16, 16, 3), np.float32)
        assert lib.rt_cuda_readback(fresh._ctx, out.ctypes.data, None) == -5          # nothing was published
        fresh.render(16, 16, -4.0, 1.0, 6)
        good, _ = fresh.readback()
        assert lib.rt_cuda_render(fresh._ctx, 16, 16, -4.0, 65536.0, 6) == -6
        again, _ = fresh.readback()                                                  # the previous frame is intact
        assert fresh.width == 16 and np.array_equal(good.view(np.uint32), again.view(np.uint32))
    finally:
        fresh.close()


def test_async_readback_overlaps_and_matches(pkg, orc_mod, oracle, gpu):
    """rt_cuda_readback_rgb8_async: a multi-frame loop with the scene resident (zoom animated), two
    readbacks in flight, pinned and pageable destinations: every frame equals the oracle's frame."""
    sph, lgt = pkg.synth_scene(48, 3, seed=6)
    W, H, S = 128, 96, 6
    zooms = [-4.0, -4.5, -5.0, -5.5, -6.0]
    gpu.upload_scene(sph, lgt)
    pinned = [pkg.HostBuffer(W * H * 3), pkg.HostBuffer(W * H * 3)]
    pageable = [np.empty(W * H * 3, np.uint8), np.empty(W * H * 3, np.uint8)]
    for bufs in (pinned, pageable):
        tickets, got = [None, None], []
        for f, z in enumerate(zooms + [None]):
            if z is not None:
                gpu.render(W, H, z, 1.0, S)
                tickets[f & 1] = gpu.readback_rgb8_async(bufs[f & 1])
            if f > 0:
                gpu.readback_wait(tickets[(f - 1) & 1])
                b = bufs[(f - 1) & 1]
                got.append(np.array(b.array if isinstance(b, pkg.HostBuffer) else b).reshape(H, W, 3).copy())
        for z, frame in zip(zooms, got):
            ref, _ = oracle.render(sph, lgt, W, H, z, 1.0, S)
            assert np.array_equal(frame, oracle.quantise(ref, oracle.max_colour(ref))), z
    for h in pinned:
        h.free()
    # the synchronous readback takes the chunked pinned path for large pageable buffers: same bytes
    sph, lgt = pkg.synth_scene(64, 4, seed=1)
    gpu.upload_scene(sph, lgt)
    gpu.render(2600, 1500, -4.0, 1.0, 4)            # 11.7 MB of RGB8: more than one 8 MiB chunk
    a = gpu.readback_rgb8()
    hb = pkg.HostBuffer(2600 * 1500 * 3)
    t = gpu.readback_rgb8_async(hb)
    gpu.readback_wait(t)
    assert np.array_equal(a.reshape(-1), hb.array)
    hb.free()


def test_filter_bound_fuzz(pkg):
    """The FMA filter of rt_core.cuh may only rule out (query, sphere) pairs the reference's exact expressions
    reject too.  > 10^9 adversarial pairs on the GPU (tests/fuzz_filter.cu: coordinates to 10^4, radii over six
    decades, rays grazing the silhouette within a relative 10^-8 .. 10^-1, unit and unnormalised directions,
    containment probes on the surface): zero violations, and the filter still rejects most pairs."""
    import ctypes
    import sys
    graft = sys.modules["__graft_entry__"]
    lib = ctypes.CDLL(str(graft.build_fuzz()))
    lib.fuzz_filter.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, ctypes.c_float,
                                ctypes.c_float, ctypes.POINTER(ctypes.c_uint64)]
    total = 0
    for seed, max_log, r_lo, r_hi in [(1, 4.0, -3.0, 3.0), (2, 1.5, -1.0, 1.0), (3, 4.0, -3.0, 0.0), (4, 3.0, 0.0, 3.0)]:
        out = (ctypes.c_uint64 * 9)()
        assert lib.fuzz_filter(seed, 1 << 18, 600, max_log, r_lo, r_hi, out) == 0
        rays, cand, hits, viol, refused, probes, pcand, inside, pviol = [int(v) for v in out]
        assert viol == 0 and pviol == 0, (seed, list(out))
        assert rays == probes == (1 << 18) * 600
        assert hits > 0.2 * rays and inside > 0.3 * probes          # the generator really aims at the surface
        assert refused < 0.01 * rays
        total += rays + probes
    assert total >= 1_000_000_000


def _needs_gpus(pkg, n):
    if pkg.device_count() < n:
        pytest.skip(f"needs {n} GPUs on this box")


def test_sin_from_cos_is_the_ieee_double_square_root(pkg):
    """rt_core.cuh sin_from_cos (a 14-instruction Markstein sequence) replaces the library's ~290-instruction IEEE double
    square root at raytracer.h:683: same float for EVERY cosine in [0, 1) — 1 065 353 216 values, checked on the GPU."""
    import ctypes
    import __graft_entry__ as graft
    lib = ctypes.CDLL(str(graft.build_fuzz()))
    lib.fuzz_sin_from_cos.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
    lib.fuzz_sin_from_cos.restype = ctypes.c_int
    out = (ctypes.c_uint64 * 2)()
    assert lib.fuzz_sin_from_cos(out) == 0
    assert out[0] == 0x3F800000 and out[1] == 0, (out[0], out[1])


@pytest.mark.parametrize("gpus", [1, 2, 4, 8])
def test_multi_gpu_c_abi_matches_oracle(pkg, orc_mod, oracle, gpus):
    """include/rt_cuda_multi.h, one process driving `gpus` devices (ncclCommInitAll): strips + NCCL max
    all-reduce + quantise + NCCL all-gather + assembly give the oracle's quantised frame byte for byte,
    with the global maximum; ragged frames (rows not divisible by strips x ranks) included."""
    _needs_gpus(pkg, gpus)
    with pkg.MultiRenderer(gpus=gpus) as m:
        assert m.world == gpus and m.local == gpus
        for (sph, lgt), (W, H, alias, S, strip) in [(pkg.default_scene(), (200, 150, 3.0, 6, 0)),
                                                    (pkg.synth_scene(300, 4, seed=4), (161, 91, 2.0, 8, 4)),
                                                    (pkg.synth_scene(64, 2, seed=8), (97, 7, 1.0, 5, 16))]:
            m.upload_scene(sph, lgt)
            for accel in (0, 2):
                m.set_option("accel", accel)
                m.render(W, H, -4.0, alias, S, strip)
                ref, ctr = oracle.render(sph, lgt, W, H, -4.0, alias, S)
                want = oracle.quantise(ref, oracle.max_colour(ref))
                for local in range(m.local):
                    frame, mx = m.readback_rgb8(local)
                    assert mx == oracle.max_colour(ref)
                    assert np.array_equal(frame, want), (gpus, local, W, H)
                assert sum(m.stats(g)["rays"] for g in range(m.local)) == ctr["rays"]
                assert m.step_ms(0) > 0
            m.set_option("accel", 0)


@pytest.mark.parametrize("gpus", [1, 2, 8])
def test_host_program_multi_gpu(pkg, tmp_path, gpus):
    """`rt_gamma --gpus N` (the C++ host over rt_cuda_multi.h, no Python in the data path) writes the
    reference CPU render's PPM."""
    import subprocess
    _needs_gpus(pkg, gpus)
    out = tmp_path / "multi.ppm"
    res = subprocess.run([str(pkg.HOST_BIN), "--gpus", str(gpus), "--out", str(out)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert hashlib.md5(out.read_bytes()).hexdigest() == FACTS["default_800x600_a3_s6"]["ppm_md5"]
    assert f"{gpus} GPUs: step" in res.stdout


def test_multi_gpu_one_process_per_gpu_nccl(pkg, tmp_path):
    """The torchrun flavour (rt_cuda_multi_init_rank, NCCL inside the library, one process per GPU) on 2
    GPUs: the assembled frame on every rank equals the oracle's quantised frame."""
    import os
    import subprocess
    import sys
    import textwrap
    _needs_gpus(pkg, 2)
    worker = tmp_path / "w.py"
    worker.write_text(textwrap.dedent("""
        import os, sys
        import numpy as np, torch, torch.distributed as dist
        sys.path.insert(0, %r)
        import __graft_entry__ as graft
        pkg = graft.load_package(); om = graft.load_oracle()
        rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        box = [pkg.multi_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        m = pkg.MultiRenderer(rank=rank, world=world, uid=box[0], device=lr)
        sph, lgt = pkg.synth_scene(200, 4, seed=3)
        W, H, alias, S = 173, 99, 2.0, 7
        m.upload_scene(sph, lgt)
        m.render(W, H, -4.0, alias, S, 4)
        frame, mx = m.readback_rgb8(0)
        oracle = om.Oracle("port")
        ref, _ = oracle.render(sph, lgt, W, H, -4.0, alias, S)
        assert mx == oracle.max_colour(ref), (mx, oracle.max_colour(ref))
        assert np.array_equal(frame, oracle.quantise(ref, mx))
        print("RANK_OK", rank)
        m.close(); dist.barrier(); dist.destroy_process_group()
    """) % str(pkg.REPO_ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29547", str(worker)],
                         capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "RANK_OK 0" in res.stdout and "RANK_OK 1" in res.stdout


def test_rare_paths(pkg, orc_mod, oracle, gpu):
    """> 4 lights (several shadow batches), candidate-list overflow, non-finite filter records,
    the largest scene the library accepts."""
    from test_hostsim import _stress_scenes
    for name, (sph, lgt) in _stress_scenes(pkg).items():
        ref, ctr = oracle.render(sph, lgt, 96, 64, -4.0, 1.0, 8)
        for opts in ({}, {"staging": 1}, {"accel": 2}):
            fb, mx, st = _render(gpu, sph, lgt, 96, 64, -4.0, 1.0, 8, **opts)
            _assert_parity(orc_mod, oracle, ref, fb)
            assert st["rays"] == ctr["rays"] and st["shadow_rays"] == ctr["shadow_rays"], (name, opts)
    sph, lgt = pkg.synth_scene(12288, 4, seed=2)          # RT_CUDA_MAX_SPHERES
    fb, mx, st = _render(gpu, sph, lgt, 48, 27, -4.0, 1.0, 8)
    ref, ctr = oracle.render(sph, lgt, 48, 27, -4.0, 1.0, 8)
    _assert_parity(orc_mod, oracle, ref, fb)
    assert st["rays"] == ctr["rays"]
    # the accelerated mode's records no longer fit one CTA at this size: the option falls back, it does not fail
    fb2, _, st2 = _render(gpu, sph, lgt, 48, 27, -4.0, 1.0, 8, accel=1)
    assert st2["accel"] == 0 and np.array_equal(orc_mod.canon(fb2), orc_mod.canon(fb))
    sph, lgt = pkg.synth_scene(6000, 4, seed=2)             # ... and still engages at 6 000
    fb, _, st = _render(gpu, sph, lgt, 48, 27, -4.0, 1.0, 8)
    fb2, _, st2 = _render(gpu, sph, lgt, 48, 27, -4.0, 1.0, 8, accel=1)
    assert st2["accel"] == 1 and np.array_equal(orc_mod.canon(fb2), orc_mod.canon(fb))


def test_strips_reassemble_to_the_full_frame(pkg, orc_mod, gpu):
    """Row-strip shards (the multi-GPU partition) tile the 1-GPU frame byte for byte,
    and the max over shards is the frame's max."""
    sph, lgt = pkg.synth_scene(64, 4, seed=1)
    W, H = 150, 101
    full, mx, _ = _render(gpu, sph, lgt, W, H, -4.0, 2.0, 6)
    for G_, strip in [(2, 8), (4, 4), (8, 16), (3, 5)]:
        asm = np.zeros_like(full)
        maxes = []
        for g in range(G_):
            gpu.render_strips(W, H, -4.0, 2.0, 6, strip, g, G_)
            part, m = gpu.readback()
            rows = pkg.local_rows(H, strip, g, G_)
            assert part.shape[0] == len(rows)
            asm[rows] = part
            maxes.append(gpu.stats()["max_colour"])
        assert np.array_equal(orc_mod.canon(asm), orc_mod.canon(full)), (G_, strip)
        assert max(maxes) == np.float32(mx)


def test_full_size_properties_config3(pkg, orc_mod, oracle, gpu):
    """BASELINE config 3 at full size (3840x2160, 256 spheres, depth 6): a fixed subset of 72
    rows spread over the frame against the oracle, plus size-independent properties."""
    sph, lgt = pkg.synth_scene(256, 4)
    W, H, S = 3840, 2160, 6
    fb, mx, st = _render(gpu, sph, lgt, W, H, -4.0, 1.0, S)
    rows = (7, 72, 30)
    ref, ctr = oracle.render(sph, lgt, W, H, -4.0, 1.0, S, rows=rows)
    got = fb[rows[0]::rows[2]][:rows[1]]
    _assert_parity(orc_mod, oracle, ref, got)
    # properties: determinism, max is the NaN-skipping max of what was written, sample count
    fb2, mx2, st2 = _render(gpu, sph, lgt, W, H, -4.0, 1.0, S)
    assert np.array_equal(orc_mod.canon(fb), orc_mod.canon(fb2)) and mx == mx2
    assert st["samples"] == W * H and st["rays"] == st2["rays"]
    assert mx == oracle.max_colour(fb)
    # quantised image: bytes equal the reference quantiser applied to the float image
    rgb = gpu.readback_rgb8()
    assert np.array_equal(rgb[rows[0]::rows[2]][:rows[1]], oracle.quantise(got, mx))


def test_host_program_writes_the_reference_ppm(pkg, tmp_path):
    """raytracer-gamma_b200/host/main.cpp (the reference's main() over the C-ABI) renders the
    default frame; the PPM it writes has the md5 of the reference CPU render's PPM."""
    import subprocess
    assert pkg.HOST_BIN.exists(), "rt_gamma was not built"
    out = tmp_path / "testPPM.ppm"
    res = subprocess.run([str(pkg.HOST_BIN), "--out", str(out)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Exec time" in res.stdout
    assert hashlib.md5(out.read_bytes()).hexdigest() == FACTS["default_800x600_a3_s6"]["ppm_md5"]
    res = subprocess.run([str(pkg.HOST_BIN), "--list"], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0 and "CUDA device" in res.stdout
    # a synthetic scene with and without the optional accelerated mode: the same file
    md5 = []
    for extra in ([], ["--accel"]):
        o = tmp_path / f"synth{len(extra)}.ppm"
        res = subprocess.run([str(pkg.HOST_BIN), "--spheres", "900", "--width", "320", "--height", "180", "--alias", "2",
                              "--depth", "8", "--out", str(o)] + extra, capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stdout + res.stderr
        md5.append(hashlib.md5(o.read_bytes()).hexdigest())
    assert md5[0] == md5[1]
    # --out x.png: the same pixels in a PNG (host/rt_png.h)
    import struct
    import zlib
    o = tmp_path / "synth.png"
    res = subprocess.run([str(pkg.HOST_BIN), "--spheres", "900", "--width", "320", "--height", "180", "--alias", "2",
                          "--depth", "8", "--out", str(o)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    data = o.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    n = struct.unpack(">I", data[33:37])[0]                      # IDAT follows the 25-byte IHDR chunk
    assert data[37:41] == b"IDAT"
    rows = np.frombuffer(zlib.decompress(data[41:41 + n]), np.uint8).reshape(180, 320 * 3 + 1)
    ppm = (tmp_path / "synth0.ppm").read_bytes()
    assert rows[:, 1:].tobytes() == ppm[ppm.index(b"255\n") + 4:]


def test_host_program_animation_and_scene_file(pkg, oracle, tmp_path):
    """`rt_gamma --frames F --zoom-step dz --scene file`: the multi-frame loop with the scene resident and
    the readback of frame f overlapping the render of frame f+1 — every frame written
    equals the oracle's frame at that zoom."""
    import subprocess
    sph, lgt = pkg.synth_scene(80, 4, seed=12)
    scene = tmp_path / "scene.txt"
    pkg.save_scene(scene, sph, lgt)
    W, H, S, F, dz = 144, 81, 7, 4, -0.75
    res = subprocess.run([str(pkg.HOST_BIN), "--scene", str(scene), "--width", str(W), "--height", str(H), "--alias", "1",
                          "--depth", str(S), "--frames", str(F), "--zoom-step", str(dz), "--out", str(tmp_path / "f_%02d.ppm")],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert f"{F} frames in" in res.stdout
    for f in range(F):
        data = (tmp_path / f"f_{f:02d}.ppm").read_bytes()
        ref, _ = oracle.render(sph, lgt, W, H, -4.0 + f * dz, 1.0, S)
        want = oracle.quantise(ref, oracle.max_colour(ref))
        assert data == b"P6\n%d %d\n255\n" % (W, H) + want.tobytes(), f


def test_reference_structs_through_the_abi(pkg, tmp_path):
    """oracle/_ref/ref_dropin (built where the reference tree exists, shipped prebuilt): the reference's
    own `struct Sphere` / `struct Light`, filled by its own setters, cast to the C-ABI's PODs — the PPM
    is the reference CPU render's, byte for byte."""
    import subprocess
    exe = pkg.REPO_ROOT / "oracle" / "_ref" / "ref_dropin"
    if not exe.exists():
        pytest.skip("oracle/_ref/ref_dropin was not built (no reference tree at build time)")
    out = tmp_path / "dropin.ppm"
    res = subprocess.run([str(exe), str(out)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert hashlib.md5(out.read_bytes()).hexdigest() == FACTS["default_800x600_a3_s6"]["ppm_md5"]


def test_full_size_properties_config4(pkg, orc_mod, oracle, gpu):
    """BASELINE config 4 at full size (7680x4320, 1024 spheres, 4 spp, depth 8): 64 rows spread
    over the frame against the oracle (the oracle needs ~0.3 s per row on 16 cores), the
    8-GPU strip partition reassembled on one GPU, and size-independent properties."""
    sph, lgt = pkg.synth_scene(1024, 4)
    W, H, S, alias = 7680, 4320, 8, 2.0
    fb, mx, st = _render(gpu, sph, lgt, W, H, -4.0, alias, S)
    rows = (33, 64, 67)
    ref, ctr = oracle.render(sph, lgt, W, H, -4.0, alias, S, rows=rows)
    got = fb[rows[0]::rows[2]][:rows[1]]
    _assert_parity(orc_mod, oracle, ref, got)
    assert st["samples"] == W * H * 4 and mx == oracle.max_colour(fb)
    # shard 5 of 8 (16-row strips) is the same pixels as the corresponding rows of the full frame
    gpu.render_strips(W, H, -4.0, alias, S, 16, 5, 8)
    part, _ = gpu.readback()
    gpu.set_option("accel", 1)
    gpu.render_strips(W, H, -4.0, alias, S, 16, 5, 8)
    part_accel, _ = gpu.readback()
    assert gpu.stats()["accel"] == 1
    gpu.set_option("accel", 0)
    assert np.array_equal(orc_mod.canon(part_accel), orc_mod.canon(part))
    mine = pkg.local_rows(H, 16, 5, 8)
    assert np.array_equal(orc_mod.canon(part), orc_mod.canon(fb[mine]))
