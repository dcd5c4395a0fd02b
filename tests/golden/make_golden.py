"""Generate the golden vectors in this directory from the REFERENCE ITSELF
(oracle/_ref/libref_s<S>.so = the reference's own headers compiled unmodified by
oracle/Makefile).  Run in the build container where /root/reference exists:

    python tests/golden/make_golden.py

The reference ships no tests or known-answer vectors for this path,
so these files are what pins the oracle — and through it the CUDA kernels — on
machines where the reference tree is absent (the GPU box).
"""
import ctypes
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
pkg.build()
om = graft.load_oracle()
om.build(reference=True)
ref = om.Oracle("reference")
port = om.Oracle("port")


def md5(b):
    return hashlib.md5(b).hexdigest()


facts = {}
sph, lgt = pkg.default_scene()

# 1. headline facts of the reference CPU render, full default frame
for S in (5, 6):
    fb, _ = ref.render(sph, lgt, 800, 600, -4.0, 3.0, S)
    mx = ref.max_colour(fb)
    q = port.quantise(fb, mx)
    ppm = b"P6\n800 600\n255\n" + q.tobytes()
    facts[f"default_800x600_a3_s{S}"] = {
        "float_md5_raw": md5(fb.tobytes()),
        "float_md5_canon": md5(om.canon(fb).tobytes()),
        "ppm_md5": md5(ppm),
        "max": float(np.float32(mx)),
        "max_bits": int(np.float32(mx).view(np.uint32)),
        "nan_pixels": int(np.isnan(fb).any(axis=2).sum()),
    }

# 2. small full framebuffers (bit patterns, NaN canonicalised)
cases = {
    "default_160x120_a2_s6": (sph, lgt, 160, 120, -4.0, 2.0, 6),
    "default_96x72_a3_s4": (sph, lgt, 96, 72, -4.0, 3.0, 4),
    "default_64x48_a1_s1": (sph, lgt, 64, 48, -4.0, 1.0, 1),
}
s256, l256 = pkg.synth_scene(256, 4)
s40, l40 = pkg.synth_scene(40, 3, seed=7)
cases["synth256_96x54_a1_s8"] = (s256, l256, 96, 54, -4.0, 1.0, 8)
cases["synth40_80x60_a2_s6"] = (s40, l40, 80, 60, -4.0, 2.0, 6)
arrays = {}
for name, (s, l, W, H, zoom, alias, S) in cases.items():
    fb, _ = ref.render(s, l, W, H, zoom, alias, S)
    arrays[name] = om.canon(fb)
    facts[name] = {"W": W, "H": H, "zoom": zoom, "alias": alias, "S": S,
                   "n": int(len(s)), "l": int(len(l)),
                   "nan_pixels": int(np.isnan(fb).any(axis=2).sum())}
arrays["synth256_spheres"] = s256.view(np.float32).reshape(-1, 12)
arrays["synth256_lights"] = l256.view(np.float32).reshape(-1, 6)
arrays["synth40_spheres"] = s40.view(np.float32).reshape(-1, 12)
arrays["synth40_lights"] = l40.view(np.float32).reshape(-1, 6)
arrays["default_spheres"] = sph.view(np.float32).reshape(-1, 12)
arrays["default_lights"] = lgt.view(np.float32).reshape(-1, 6)

# 3. known-answer vectors of the unit functions, from the reference
rng = np.random.default_rng(20261018)
lib = ref._ref(6)
K = 512
kat_sph = np.zeros(K, pkg.SPHERE_DTYPE)
kat_sph["pos"] = rng.uniform(-10, 10, (K, 3)).astype(np.float32)
kat_sph["radius"] = rng.uniform(0.1, 4, K).astype(np.float32)
o = rng.uniform(-12, 12, (K, 3)).astype(np.float32)
d = rng.normal(size=(K, 3)).astype(np.float32)
d[::7] /= np.linalg.norm(d[::7], axis=1, keepdims=True).astype(np.float32)
# aim most rays at their sphere so that hits, grazes and misses all occur
aim = kat_sph["pos"] + rng.normal(scale=1.0, size=(K, 3)).astype(np.float32) * kat_sph["radius"][:, None]
d[1::2] = (aim - o)[1::2]
d[5] = 0.0                                  # the degenerate ray of total internal reflection
o[9] = kat_sph["pos"][9] + np.float32([kat_sph["radius"][9], 0, 0])   # origin on the surface
hit = np.zeros(K, np.int32)
tval = np.zeros(K, np.float32)
for i in range(K):
    t = ctypes.c_float(0)
    hit[i] = lib.ref_ray_sphere(kat_sph[i:i + 1].ctypes.data, o[i].ctypes.data, d[i].ctypes.data,
                                ctypes.addressof(t))
    tval[i] = t.value if hit[i] else 0.0
arrays.update(kat_rs_spheres=kat_sph.view(np.float32).reshape(-1, 12), kat_rs_o=o, kat_rs_d=d,
              kat_rs_hit=hit, kat_rs_t=tval.view(np.uint32))
facts["kat_ray_sphere"] = {"n": K, "hits": int(hit.sum())}

pts = rng.uniform(-12, 12, (K, 3)).astype(np.float32)
pts[::3] = (s40["pos"][rng.integers(0, 40, len(pts[::3]))]
            + rng.normal(scale=0.5, size=(len(pts[::3]), 3)).astype(np.float32))
cont = np.array([lib.ref_primary_container(s40.ctypes.data, 40, pts[i].ctypes.data) for i in range(K)], np.int32)
arrays.update(kat_pc_pts=pts, kat_pc_idx=cont)
facts["kat_primary_container"] = {"n": K, "inside": int((cont >= 0).sum())}

abc = rng.uniform(-3, 3, (K, 3)).astype(np.float32)
abc[::5, 0] = rng.uniform(-0.002, 0.002, len(abc[::5])).astype(np.float32)
abc[::11, 1] = rng.uniform(-0.002, 0.002, len(abc[::11])).astype(np.float32)
nroots = np.zeros(K, np.int32)
roots = np.zeros((K, 2), np.float32)
for i in range(K):
    r = (ctypes.c_float * 2)(0, 0)
    nroots[i] = lib.ref_solve_quadratic(abc[i, 0], abc[i, 1], abc[i, 2], ctypes.addressof(r))
    roots[i, :nroots[i]] = [r[j] for j in range(nroots[i])]
arrays.update(kat_sq_abc=abc, kat_sq_n=nroots, kat_sq_roots=om.canon(roots))

# materials through the reference's setters (raytracer.h:59-74)
mats = rng.uniform(0, 1, (64, 9)).astype(np.float32)
mat_out = np.zeros((64, 8), np.float32)
for i in range(64):
    lib.ref_make_material(mat_out[i].ctypes.data, mats[i, 0:3].ctypes.data, mats[i, 3:6].ctypes.data,
                          mats[i, 6], mats[i, 7], np.float32(1 + mats[i, 8]))
arrays.update(kat_mat_in=mats, kat_mat_out=mat_out)

np.savez_compressed(HERE / "golden.npz", **arrays)
(HERE / "facts.json").write_text(json.dumps(facts, indent=1, sort_keys=True) + "\n")
print(json.dumps(facts, indent=1, sort_keys=True))
print("wrote", HERE / "golden.npz", (HERE / "golden.npz").stat().st_size, "bytes")
