"""CPU tests: the oracle (C restatement) against the golden vectors generated from the
reference itself (tests/golden/make_golden.py) and, where /root/reference is present,
against the reference harness directly."""
import ctypes
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"
FACTS = json.loads((GOLD / "facts.json").read_text())
G = np.load(GOLD / "golden.npz")


def _scene(pkg, name):
    sph = np.ascontiguousarray(G[f"{name}_spheres"]).view(pkg.SPHERE_DTYPE).reshape(-1)
    lgt = np.ascontiguousarray(G[f"{name}_lights"]).view(pkg.LIGHT_DTYPE).reshape(-1)
    return sph, lgt


def _case_scene(pkg, case):
    return _scene(pkg, case.split("_")[0])


FB_CASES = [k for k in FACTS if "W" in FACTS[k]]


def test_scene_builders_match_golden(pkg):
    sph, lgt = pkg.default_scene()
    assert np.array_equal(sph.view(np.float32).reshape(-1, 12), G["default_spheres"])
    assert np.array_equal(lgt.view(np.float32).reshape(-1, 6), G["default_lights"])
    s, l = pkg.synth_scene(256, 4)
    assert np.array_equal(s.view(np.float32).reshape(-1, 12), G["synth256_spheres"])
    assert np.array_equal(l.view(np.float32).reshape(-1, 6), G["synth256_lights"])
    s, l = pkg.synth_scene(40, 3, seed=7)
    assert np.array_equal(s.view(np.float32).reshape(-1, 12), G["synth40_spheres"])


def test_make_material_matches_reference_setters(pkg):
    m = G["kat_mat_in"]
    for i in range(len(m)):
        got = pkg.make_material(m[i, 0:3], m[i, 3:6], float(m[i, 6]), float(m[i, 7]), float(np.float32(1 + m[i, 8])))
        assert np.array_equal(got.view(np.uint32), G["kat_mat_out"][i].view(np.uint32)), i


def test_synth_scene_is_well_formed(pkg):
    for n in (16, 1024, 4096):
        s, l = pkg.synth_scene(n, 4)
        assert len(s) == n and len(l) == 4
        assert np.all(s["radius"] > 0) and np.all(s["pos"][:, 2] <= -6) and np.all(s["pos"][:, 2] >= -24)
        assert np.all((s["opacity"] > 0) & (s["opacity"] <= 1))
        if n >= 1024:
            assert 0.2 < (s["opacity"] == 1).mean() < 0.3      # a quarter are opaque
        assert np.all(np.abs(s["pos"][:, 0]) <= 0.9 * 2.667 * 24 + 1e-3)
    a, _ = pkg.synth_scene(64, 4, seed=1)
    b, _ = pkg.synth_scene(64, 4, seed=2)
    assert not np.array_equal(a.view(np.float32), b.view(np.float32))


@pytest.mark.parametrize("S", [5, 6])
def test_default_frame_facts(pkg, orc_mod, oracle, S):
    """max 6.95503e-05, 2492 NaN pixels at S=6, none at S=5, PPM md5."""
    f = FACTS[f"default_800x600_a3_s{S}"]
    sph, lgt = pkg.default_scene()
    fb, ctr = oracle.render(sph, lgt, 800, 600, -4.0, 3.0, S)
    assert hashlib.md5(orc_mod.canon(fb).tobytes()).hexdigest() == f["float_md5_canon"]
    mx = oracle.max_colour(fb)
    assert int(np.float32(mx).view(np.uint32)) == f["max_bits"]
    assert int(np.isnan(fb).any(axis=2).sum()) == f["nan_pixels"]
    ppm = b"P6\n800 600\n255\n" + oracle.quantise(fb, mx).tobytes()
    assert hashlib.md5(ppm).hexdigest() == f["ppm_md5"]
    if S == 6:
        assert f["nan_pixels"] == 2492 and abs(mx - 6.95503e-05) < 1e-10
        assert f["float_md5_raw"] == "96b20332f1c1fdf97efec7da14a0d920"
        assert f["ppm_md5"] == "77a498a83918ef392d1f080532d1f4dc"
        # work counters of the reference algorithm (SURVEY.md §3.3 probe)
        assert ctr["rays"] == 16617984 and ctr["shadow_rays"] == 6928426
        assert ctr["sphere_tests"] == 49853952 and ctr["dropped_pushes"] == 808106
        assert ctr["max_stack"] == 6 and ctr["samples"] == 4320000
    else:
        assert f["nan_pixels"] == 0


@pytest.mark.parametrize("case", FB_CASES)
def test_oracle_matches_golden_framebuffers(pkg, orc_mod, oracle, case):
    f = FACTS[case]
    sph, lgt = _case_scene(pkg, case)
    fb, _ = oracle.render(sph, lgt, f["W"], f["H"], f["zoom"], f["alias"], f["S"])
    assert np.array_equal(orc_mod.canon(fb), G[case])
    assert int(np.isnan(fb).any(axis=2).sum()) == f["nan_pixels"]


def test_kat_ray_sphere(pkg, oracle):
    sph = np.ascontiguousarray(G["kat_rs_spheres"]).view(pkg.SPHERE_DTYPE).reshape(-1)
    o, d = np.ascontiguousarray(G["kat_rs_o"]), np.ascontiguousarray(G["kat_rs_d"])
    hits = 0
    for i in range(len(sph)):
        t = ctypes.c_float(0)
        h = oracle._port.rt_oracle_ray_sphere(sph[i:i + 1].ctypes.data, o[i].ctypes.data, d[i].ctypes.data,
                                              ctypes.addressof(t))
        assert h == G["kat_rs_hit"][i], i
        if h:
            assert np.float32(t.value).view(np.uint32) == G["kat_rs_t"][i], i
            hits += 1
    assert hits == FACTS["kat_ray_sphere"]["hits"]
    assert G["kat_rs_hit"][5] == 0      # zero direction never hits (0/0 = NaN roots)


def test_kat_primary_container(pkg, oracle):
    sph, _ = _scene(pkg, "synth40")
    pts = np.ascontiguousarray(G["kat_pc_pts"])
    got = [oracle._port.rt_oracle_primary_container(sph.ctypes.data, 40, pts[i].ctypes.data) for i in range(len(pts))]
    assert np.array_equal(np.array(got, np.int32), G["kat_pc_idx"])


def test_kat_solve_quadratic(orc_mod, oracle):
    abc = G["kat_sq_abc"]
    for i in range(len(abc)):
        r = (ctypes.c_float * 2)(0, 0)
        n = oracle._port.rt_oracle_solve_quadratic(abc[i, 0], abc[i, 1], abc[i, 2], ctypes.addressof(r))
        assert n == G["kat_sq_n"][i]
        got = np.zeros(2, np.float32)
        got[:n] = [r[j] for j in range(n)]
        assert np.array_equal(orc_mod.canon(got), G["kat_sq_roots"][i]), i


def test_quantiser_edge_cases(oracle):
    """main.cpp:71-76 incl. the x86-64 behaviour of the undefined casts."""
    mx = 6.955025310162455e-05
    fb = np.array([[0.0, mx, mx / 2], [np.nan, -1.0, 2.0], [1e-7, 1.0, -0.0]], np.float32).reshape(1, 3, 3)
    q = oracle.quantise(fb, mx).reshape(-1)
    m32 = np.float32(mx)

    def expect(v):   # float32 arithmetic in the reference's order: (min(1,v) * 255) / max, truncated
        return int(np.float32(np.float32(min(1.0, v)) * np.float32(255)) / m32) & 0xFF

    assert q[0] == 0 and q[1] == expect(float(fb[0, 0, 1])) and q[1] >= 254 and q[2] == 127
    assert q[3] == 237                    # NaN -> min(1,NaN)=1 -> 255/max -> low byte (SURVEY hard part 2)
    assert q[5] == 237 and q[7] == 237    # anything >= 1 clips to the same value
    assert q[4] == (int(-1.0 * 255 / np.float32(mx)) & 0xFF)
    # a tiny max pushes 255/max past INT_MAX: cvttss2si gives INT_MIN, low byte 0
    assert oracle.quantise(np.full((1, 1, 3), np.nan, np.float32), 1e-9).reshape(-1)[0] == 0
    assert oracle.max_colour(np.zeros((2, 2, 3), np.float32)) == 1.0
    assert oracle.max_colour(np.array([[np.nan, 0.5, -3.0]], np.float32)) == 0.5


def test_row_subset_equals_full_frame(pkg, orc_mod, oracle):
    sph, lgt = pkg.default_scene()
    full, _ = oracle.render(sph, lgt, 64, 48, -4.0, 2.0, 6)
    part, _ = oracle.render(sph, lgt, 64, 48, -4.0, 2.0, 6, rows=(3, 9, 5))
    assert np.array_equal(orc_mod.canon(part), orc_mod.canon(full[3::5][:9]))


@pytest.mark.parametrize("S", [1, 2, 3, 4, 5, 6, 7, 8])
def test_port_is_bit_identical_to_reference(pkg, orc_mod, oracle, reference, S):
    """The pin: rt_oracle.c against the reference's own headers, every stack size."""
    if not orc_mod.reference_available(S):
        pytest.skip("this stack size was not built")
    for (sph, lgt), (W, H, alias) in [(pkg.default_scene(), (120, 90, 2.0)),
                                      (pkg.synth_scene(48, 4, seed=S), (72, 40, 1.0))]:
        a, _ = oracle.render(sph, lgt, W, H, -4.0, alias, S)
        b, _ = reference.render(sph, lgt, W, H, -4.0, alias, S)
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.array_equal(orc_mod.canon(a), orc_mod.canon(b))
        assert oracle.max_colour(a) == reference.max_colour(b)


def test_port_matches_reference_odd_parameters(pkg, orc_mod, oracle, reference):
    sph, lgt = pkg.synth_scene(24, 2, seed=3)
    for W, H, zoom, alias, S in [(33, 17, -4.0, 1.0, 6), (50, 40, -2.5, 2.5, 6), (16, 16, -6.0, 0.5, 6),
                                 (40, 30, -4.0, 3.0, 5)]:
        a, _ = oracle.render(sph, lgt, W, H, zoom, alias, S)
        b, _ = reference.render(sph, lgt, W, H, zoom, alias, S)
        assert np.array_equal(orc_mod.canon(a), orc_mod.canon(b)), (W, H, zoom, alias, S)
    # no spheres / no lights
    a, _ = oracle.render(sph[:0], lgt, 16, 12, -4.0, 1.0, 6)
    b, _ = reference.render(sph[:0], lgt, 16, 12, -4.0, 1.0, 6)
    assert np.array_equal(a, b) and not a.any()
    a, _ = oracle.render(sph, lgt[:0], 16, 12, -4.0, 1.0, 6)
    b, _ = reference.render(sph, lgt[:0], 16, 12, -4.0, 1.0, 6)
    assert np.array_equal(orc_mod.canon(a), orc_mod.canon(b))
