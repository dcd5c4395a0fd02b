"""CPU tests of the kernel's state machine (rt_core.cuh) through the lane simulator
(tests/hostsim.cpp): same machine, same filter, same exact tests as the CUDA kernel,
run one lane at a time and compared bit-for-bit with the oracle."""
import json
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"
FACTS = json.loads((GOLD / "facts.json").read_text())
G = np.load(GOLD / "golden.npz")
FB_CASES = [k for k in FACTS if "W" in FACTS[k]]


def _case_scene(pkg, case):
    name = case.split("_")[0]
    sph = np.ascontiguousarray(G[f"{name}_spheres"]).view(pkg.SPHERE_DTYPE).reshape(-1)
    lgt = np.ascontiguousarray(G[f"{name}_lights"]).view(pkg.LIGHT_DTYPE).reshape(-1)
    return sph, lgt


@pytest.mark.parametrize("case", FB_CASES)
def test_machine_matches_golden(pkg, orc_mod, hostsim, case):
    f = FACTS[case]
    sph, lgt = _case_scene(pkg, case)
    fb, _ = hostsim(sph, lgt, f["W"], f["H"], f["zoom"], f["alias"], f["S"])
    assert np.array_equal(orc_mod.canon(fb), G[case])


@pytest.mark.parametrize("S", [1, 2, 5, 6, 8, 12])
def test_machine_matches_oracle_counters_and_pixels(pkg, orc_mod, oracle, hostsim, S):
    """Work counters are the reference algorithm's: rays, shadow rays, container queries/tests."""
    sph, lgt = pkg.default_scene()
    a, ca = oracle.render(sph, lgt, 200, 150, -4.0, 3.0, S)
    b, cb = hostsim(sph, lgt, 200, 150, -4.0, 3.0, S)
    assert np.array_equal(orc_mod.canon(a), orc_mod.canon(b))
    for k in ("rays", "shadow_rays", "contain_queries", "contain_tests", "samples"):
        assert ca[k] == cb[k], k


def test_filter_never_changes_the_image(pkg, orc_mod, hostsim):
    """The FMA filter only removes certain misses: with it off (exact test against every
    sphere) the framebuffer is the same, and far fewer exact tests are needed with it on."""
    for n, seed in [(64, 0), (300, 5)]:
        sph, lgt = pkg.synth_scene(n, 4, seed=seed)
        a, ca = hostsim(sph, lgt, 96, 64, -4.0, 1.0, 8)
        b, cb = hostsim(sph, lgt, 96, 64, -4.0, 1.0, 8, no_filter=True)
        assert np.array_equal(orc_mod.canon(a), orc_mod.canon(b))
        assert ca["rays"] == cb["rays"]
        assert ca["exact_tests"] < cb["exact_tests"] / 20


def test_machine_matches_oracle_synthetic(pkg, orc_mod, oracle, hostsim):
    for n, l, seed, W, H, alias, S in [(256, 4, 0, 120, 68, 2.0, 8), (1024, 4, 0, 64, 36, 1.0, 8),
                                       (17, 1, 9, 57, 33, 1.0, 4), (33, 0, 2, 40, 30, 1.0, 6)]:
        sph, lgt = pkg.synth_scene(n, l, seed=seed)
        a, ca = oracle.render(sph, lgt, W, H, -4.0, alias, S)
        b, cb = hostsim(sph, lgt, W, H, -4.0, alias, S)
        assert np.array_equal(orc_mod.canon(a), orc_mod.canon(b)), (n, l, seed)
        assert ca["rays"] == cb["rays"] and ca["contain_tests"] == cb["contain_tests"]


def test_machine_edge_cases(pkg, orc_mod, oracle, hostsim):
    sph, lgt = pkg.default_scene()
    # fractional / sub-unit alias factors change the sample count (int i < float alias)
    for alias in (0.5, 1.5, 2.5):
        a, _ = oracle.render(sph, lgt, 40, 30, -4.0, alias, 6)
        b, _ = hostsim(sph, lgt, 40, 30, -4.0, alias, 6)
        assert np.array_equal(orc_mod.canon(a), orc_mod.canon(b)), alias
    # empty scene: every ray misses, the image is black
    b, cb = hostsim(sph[:0], lgt, 16, 12, -4.0, 1.0, 6)
    assert not b.any() and cb["rays"] == 16 * 12
    # far spheres beyond kMaxRenderDist = 1000 are ignored (raytracer.h:156)
    far = sph.copy()
    far["pos"][:, 2] -= 1200
    a, _ = oracle.render(far, lgt, 32, 24, -4.0, 1.0, 6)
    b, _ = hostsim(far, lgt, 32, 24, -4.0, 1.0, 6)
    assert np.array_equal(a, b) and not a.any()
    # camera inside a sphere, overlapping spheres (container order matters, raytracer.h:264)
    inside = sph.copy()
    inside["pos"][0] = (0, 0, -1)
    inside["radius"][0] = 3
    inside["pos"][1] = (0.5, 0, -2)
    a, _ = oracle.render(inside, lgt, 48, 36, -4.0, 1.0, 6)
    b, _ = hostsim(inside, lgt, 48, 36, -4.0, 1.0, 6)
    assert np.array_equal(orc_mod.canon(a), orc_mod.canon(b))


def _stress_scenes(pkg):
    """Scenes that push the kernel's rare paths: > 4 lights (several shadow batches per hit),
    many spheres pierced by one ray (candidate-list overflow -> exact fallback), huge
    coordinates (filter values overflow -> non-filterable geometry -> exact fallback)."""
    out = {}
    sph, lgt = pkg.synth_scene(48, 4, seed=21)
    six = np.zeros(6, pkg.LIGHT_DTYPE)
    six[:4] = lgt
    six[4]["pos"], six[4]["col"] = (5, 90, 30), (0.4, 0.2, 0.6)
    six[5]["pos"], six[5]["col"] = (-60, -40, 10), (0.3, 0.3, 0.1)
    out["six_lights"] = (sph, six)
    # 40 thin transparent shells stacked along the view axis
    shells = np.zeros(40, pkg.SPHERE_DTYPE)
    for i in range(40):
        shells[i]["pos"] = (0.05 * i, 0.0, -8.0 - 0.01 * i)
        shells[i]["radius"] = 2.0 + 0.05 * i
        shells[i]["matte"] = (0.3, 0.5, 0.7)
        shells[i]["gloss"] = (0.1, 0.1, 0.1)
        shells[i]["opacity"] = 0.35
        shells[i]["refractiveIndex"] = 1.0 + 0.01 * i
    out["shells"] = (shells, lgt[:2])
    far = sph.copy()
    far["pos"][5] = (3e19, 1e19, -2e19)       # |c|^2 overflows float: its filter record is not finite
    far["radius"][5] = 1e19
    out["huge_sphere"] = (far, lgt)
    return out


def test_machine_rare_paths(pkg, orc_mod, oracle, hostsim):
    for name, (sph, lgt) in _stress_scenes(pkg).items():
        a, ca = oracle.render(sph, lgt, 64, 48, -4.0, 1.0, 8)
        b, cb = hostsim(sph, lgt, 64, 48, -4.0, 1.0, 8)
        assert np.array_equal(orc_mod.canon(a), orc_mod.canon(b)), name
        assert ca["rays"] == cb["rays"] and ca["shadow_rays"] == cb["shadow_rays"], name


def test_cluster_filter_is_conservative_and_exact(pkg, orc_mod, oracle, hostsim):
    """Accelerated mode (two-level cluster filter): same pixels as the oracle, and an audit of EVERY
    query finds no sphere the exact test accepts inside a cluster the cluster filter ruled out."""
    for n, l, seed, W, H, alias, S in [(256, 4, 0, 96, 54, 2.0, 8), (1024, 4, 0, 64, 36, 1.0, 8),
                                       (100, 3, 7, 64, 48, 1.0, 6), (37, 1, 9, 57, 33, 1.0, 4)]:
        sph, lgt = pkg.synth_scene(n, l, seed=seed)
        a, ca = oracle.render(sph, lgt, W, H, -4.0, alias, S)
        b, cb = hostsim(sph, lgt, W, H, -4.0, alias, S, mode=3)
        assert np.array_equal(orc_mod.canon(a), orc_mod.canon(b)), (n, l, seed)
        assert cb["accel_violations"] == 0
        assert ca["rays"] == cb["rays"] and ca["shadow_rays"] == cb["shadow_rays"]
        queries = cb["rays"] + cb["contain_queries"]
        assert cb["cluster_tests"] <= queries * ((n + 7) // 8 + 1)      # an eighth of the brute-force filter tests


def test_cluster_filter_rare_paths(pkg, orc_mod, oracle, hostsim):
    """The stress scenes (nested / coincident / huge / tiny / non-finite spheres, light inside a sphere)
    through the accelerated mode, audited."""
    for name, (sph, lgt) in _stress_scenes(pkg).items():
        a, ca = oracle.render(sph, lgt, 64, 48, -4.0, 1.0, 8)
        b, cb = hostsim(sph, lgt, 64, 48, -4.0, 1.0, 8, mode=3)
        assert np.array_equal(orc_mod.canon(a), orc_mod.canon(b)), name
        assert cb["accel_violations"] == 0, name
    # far-away, large-coordinate scene: the slack terms scale with |o|^2 and |c|^2
    sph, lgt = pkg.synth_scene(200, 2, seed=3)
    far = sph.copy()
    far["pos"] = far["pos"] * np.float32(37.0)
    far["radius"] = far["radius"] * np.float32(37.0)
    flt = lgt.copy()
    flt["pos"] = flt["pos"] * np.float32(37.0)
    a, _ = oracle.render(far, flt, 64, 48, -4.0, 1.0, 8)
    b, cb = hostsim(far, flt, 64, 48, -4.0, 1.0, 8, mode=3)
    assert np.array_equal(orc_mod.canon(a), orc_mod.canon(b)) and cb["accel_violations"] == 0


def test_cluster_builder_partitions_and_bounds(pkg):
    """build_clusters (rt_soa.h): every sphere is a member of exactly one cluster, clusters hold <= 8,
    the bounding record's radius (recovered from w = |C|^2 - R^2 - kappa(|C|^2 + R^2) - ...) reaches past
    every member, spheres the filter cannot represent sit in always-candidate clusters."""
    import ctypes
    import __graft_entry__ as graft
    lib = ctypes.CDLL(str(graft.build_hostsim()))
    lib.hostsim_clusters.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint]
    lib.hostsim_clusters.restype = ctypes.c_int
    kappa = 2.0 ** -17
    for n, seed, odd in [(1, 0, 0), (8, 1, 0), (9, 2, 0), (100, 3, 0), (1024, 0, 0), (777, 5, 13)]:
        sph, _ = pkg.synth_scene(n, 1, seed=seed)
        sph = sph.copy()
        for k in range(odd):                                   # NaN / infinite geometry
            sph["pos"][3 + 7 * k] = (np.float32("nan"), 0.0, 0.0) if k % 2 else (np.float32("inf"), 1.0, 2.0)
        cap = n // 8 + 4
        rec = np.zeros((cap, 4), np.float32)
        idx = np.zeros((cap, 8), np.uint16)
        nc = lib.hostsim_clusters(sph.ctypes.data, n, rec.ctypes.data, idx.ctypes.data, cap)
        assert 0 < nc <= (n + 7) // 8 + 1
        members = idx[:nc].reshape(-1)
        real = members[members != 0x3FFF]
        assert np.array_equal(np.sort(real), np.arange(n))       # a partition of the spheres
        finite = np.isfinite(sph["pos"]).all(axis=1) & np.isfinite(sph["radius"])
        for c in range(nc):
            m = idx[c][idx[c] != 0x3FFF]
            assert 1 <= len(m) <= 8
            if rec[c, 3] == -np.inf:                             # always a candidate
                assert not finite[m].any()
                continue
            assert finite[m].all()
            C = rec[c, :3].astype(np.float64)
            cc = float(C @ C)
            # w <= cc - R^2 - kappa (cc + R^2)  =>  R^2 >= (cc (1 - kappa) - w) / (1 + kappa) - (small terms)
            R = np.sqrt(max(0.0, (cc * (1 - kappa) - float(rec[c, 3])) / (1 + kappa)))
            reach = np.linalg.norm(sph["pos"][m].astype(np.float64) - C, axis=1) + np.abs(sph["radius"][m].astype(np.float64))
            assert (reach <= R * (1 + 1e-6)).all(), (n, c, reach.max(), R)
            assert R <= 1.25 * reach.max() + 0.3                 # and is not wildly loose


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_cluster_filter_audit_under_rescaling(pkg, orc_mod, oracle, hostsim, seed):
    """The cluster bound's slack terms scale with |o|^2, |c|^2 and r^2: audit translated, shrunk and
    blown-up copies of random scenes (and mixed radii spanning four decades) — no exactly-accepted sphere
    may sit in a ruled-out cluster, and the frame stays the oracle's."""
    rng = np.random.default_rng(seed)
    sph, lgt = pkg.synth_scene(int(rng.integers(40, 400)), int(rng.integers(1, 5)), seed=seed)
    variants = []
    for scale in (np.float32(0.01), np.float32(1.0), np.float32(300.0)):
        s, l = sph.copy(), lgt.copy()
        s["pos"] *= scale; s["radius"] *= scale; l["pos"] *= scale
        variants.append((f"scale {scale}", s, l))
    s = sph.copy()
    s["radius"] *= (10.0 ** rng.uniform(-2.5, 1.0, len(s))).astype(np.float32)      # tiny next to huge
    variants.append(("mixed radii", s, lgt))
    s = sph.copy()
    s["pos"][:, 2] -= np.float32(2000.0)                                             # far from the camera
    variants.append(("far away", s, lgt))
    s = sph.copy()
    s["pos"][: len(s) // 2] = s["pos"][0]                                            # half of them concentric
    variants.append(("concentric", s, lgt))
    for name, s, l in variants:
        a, _ = oracle.render(s, l, 48, 36, -4.0, 1.0, 8)
        b, cb = hostsim(s, l, 48, 36, -4.0, 1.0, 8, mode=3)
        assert cb["accel_violations"] == 0, (seed, name)
        assert np.array_equal(orc_mod.canon(a), orc_mod.canon(b)), (seed, name)


@pytest.mark.parametrize("W,H,alias,strip", [(64, 48, 1.0, (48, 0, 1)), (61, 37, 1.0, (37, 0, 1)), (61, 37, 2.0, (37, 0, 1)),
                                             (200, 150, 3.0, (16, 1, 3)), (33, 70, 1.0, (4, 2, 8)), (130, 9, 1.0, (4, 0, 2)),
                                             (7, 3, 1.0, (3, 0, 1)), (40, 40, 5.0, (4, 3, 4))])
def test_work_map_covers_every_sample_once(W, H, alias, strip):
    """rt_core.cuh WorkMap: groups of 32 items (8x4 tiles x samples, or 16x8 tiles x four pixel sub-lattices at
    1 spp) produce every result record of the shard exactly once, on this shard's rows, and tile_of_dst inverts it."""
    import ctypes
    import __graft_entry__ as graft
    lib = ctypes.CDLL(str(graft.build_hostsim()))
    lib.hostsim_workmap.argtypes = [ctypes.c_uint] * 7 + [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint)]
    lib.hostsim_workmap.restype = ctypes.c_int
    n_iter = int(np.ceil(alias))
    spp = n_iter * n_iter
    rows, first, stride = strip
    local = ctypes.c_uint(0)
    hits = np.zeros(W * H * spp, np.uint32)            # at least localRows * W * spp
    bad = lib.hostsim_workmap(W, H, rows, first, stride, spp, n_iter, hits.ctypes.data, ctypes.byref(local))
    assert bad == 0
    n = local.value * W * spp
    assert n > 0 and (hits[:n] == 1).all() and (hits[n:] == 0).all()
