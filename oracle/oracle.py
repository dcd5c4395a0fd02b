"""ctypes binding of the CPU oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (raytracer-gamma_b200)
never does.

  Oracle("port")        oracle/_build/librt_oracle.so  (rt_oracle.c, the C restatement)
  Oracle("reference")   oracle/_ref/libref_s<S>.so     (the reference's own headers, one build per stack size)
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PORT_LIB = HERE / "_build" / "librt_oracle.so"
REF_DIR = HERE / "_ref"
REFERENCE_TREE = Path("/root/reference/raytracer_gamma")

COUNTER_NAMES = ["rays", "shadow_rays", "sphere_tests", "contain_queries", "contain_tests",
                 "refractions", "reflections", "pops", "dropped_pushes", "samples", "max_stack"]

_RENDER_ARGS = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint,
                ctypes.c_uint, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_uint,
                ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]


def build(reference: bool = True) -> None:
    """Build the port, and the reference harness when /root/reference is present."""
    targets = ["oracle"] + (["ref"] if reference and REFERENCE_TREE.is_dir() else [])
    subprocess.run(["make", "-s", "-C", str(HERE), *targets], check=True)


def reference_available(stack: int) -> bool:
    return (REF_DIR / f"libref_s{stack}.so").exists()


class Oracle:
    def __init__(self, kind: str = "port"):
        assert kind in ("port", "reference")
        self.kind = kind
        self._libs: dict = {}
        if kind == "port":
            if not PORT_LIB.exists():
                build(reference=False)
            lib = ctypes.CDLL(str(PORT_LIB))
            lib.rt_oracle_render.argtypes = _RENDER_ARGS
            lib.rt_oracle_render.restype = ctypes.c_int
            lib.rt_oracle_max_colour.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
            lib.rt_oracle_max_colour.restype = ctypes.c_float
            lib.rt_oracle_quantise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float, ctypes.c_void_p]
            lib.rt_oracle_ray_sphere.argtypes = [ctypes.c_void_p] * 4
            lib.rt_oracle_primary_container.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p]
            lib.rt_oracle_solve_quadratic.argtypes = [ctypes.c_float] * 3 + [ctypes.c_void_p]
            lib.rt_oracle_closest_hit.argtypes = [ctypes.c_void_p, ctypes.c_uint] + [ctypes.c_void_p] * 5
            self._port = lib

    def _ref(self, stack: int):
        if stack not in self._libs:
            path = REF_DIR / f"libref_s{stack}.so"
            if not path.exists():
                raise FileNotFoundError(f"{path} missing (built by `make -C oracle ref` where /root/reference exists)")
            lib = ctypes.CDLL(str(path))
            lib.ref_render.argtypes = _RENDER_ARGS
            lib.ref_render.restype = ctypes.c_int
            lib.ref_max_colour.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
            lib.ref_max_colour.restype = ctypes.c_float
            lib.ref_ray_sphere.argtypes = [ctypes.c_void_p] * 4
            lib.ref_primary_container.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p]
            lib.ref_solve_quadratic.argtypes = [ctypes.c_float] * 3 + [ctypes.c_void_p]
            lib.ref_make_material.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_float] * 3
            self._libs[stack] = lib
        return self._libs[stack]

    def render(self, spheres, lights, width, height, zoom=-4.0, alias=1.0, max_stack=6,
               rows=None, threads=0):
        """rows: None (whole frame) or (begin, count, step).  -> (float32 [count, W, 3], counters dict)"""
        spheres = np.ascontiguousarray(spheres)
        lights = np.ascontiguousarray(lights)
        begin, count, step = (0, height, 1) if rows is None else rows
        out = np.zeros((count, width, 3), np.float32)
        ctr = (ctypes.c_uint64 * len(COUNTER_NAMES))()
        fn = self._port.rt_oracle_render if self.kind == "port" else self._ref(max_stack).ref_render
        rc = fn(spheres.ctypes.data if len(spheres) else None, len(spheres),
                lights.ctypes.data if len(lights) else None, len(lights),
                width, height, zoom, alias, max_stack, begin, count, step,
                out.ctypes.data, ctypes.addressof(ctr), threads)
        if rc:
            raise RuntimeError(f"oracle render failed: {rc}")
        return out, dict(zip(COUNTER_NAMES, [int(v) for v in ctr]))

    def max_colour(self, fb: np.ndarray) -> float:
        fb = np.ascontiguousarray(fb, np.float32)
        if self.kind == "port":
            return float(self._port.rt_oracle_max_colour(fb.ctypes.data, fb.size // 3))
        return float(self._ref(6).ref_max_colour(fb.ctypes.data, fb.size // 3))

    def quantise(self, fb: np.ndarray, max_colour: float) -> np.ndarray:
        """main.cpp:71-76 — always the port (savePPM cannot be compiled here: main.cpp needs OpenCL)."""
        fb = np.ascontiguousarray(fb, np.float32)
        out = np.zeros(fb.shape, np.uint8)
        lib = self._port if self.kind == "port" else Oracle("port")._port
        lib.rt_oracle_quantise(fb.ctypes.data, fb.size // 3, max_colour, out.ctypes.data)
        return out


def canon(fb: np.ndarray) -> np.ndarray:
    """Bit pattern of a float image with every NaN mapped to one value: NaN payload/sign
    bits are not part of the reference's semantics (x86 propagates operand payloads,
    the GPU returns the canonical NaN), NaN-ness is."""
    bits = np.ascontiguousarray(fb, np.float32).view(np.uint32).copy()
    bits[np.isnan(fb)] = 0x7FC00000
    return bits


def compare(ref_fb: np.ndarray, got_fb: np.ndarray, oracle: "Oracle | None" = None) -> dict:
    """Parity report per SURVEY.md §8(d): NaN masks, bit equality, and the 8-bit
    comparison after the reference quantiser with the ORACLE's max."""
    o = oracle or Oracle("port")
    ref_fb = np.ascontiguousarray(ref_fb, np.float32)
    got_fb = np.ascontiguousarray(got_fb, np.float32)
    nan_equal = bool(np.array_equal(np.isnan(ref_fb), np.isnan(got_fb)))
    bit_diff_px = int((canon(ref_fb) != canon(got_fb)).any(axis=-1).sum())
    mx = o.max_colour(ref_fb)
    a = o.quantise(ref_fb, mx).astype(np.int16)
    b = o.quantise(got_fb, mx).astype(np.int16)
    d = np.abs(a - b).max(axis=-1)
    npx = d.size
    off = np.argwhere(d > 1)
    return {
        "pixels": int(npx), "nan_masks_equal": nan_equal, "bit_different_pixels": bit_diff_px,
        "bit_exact": nan_equal and bit_diff_px == 0,
        "within_1lsb_frac": float((d <= 1).sum() / max(1, npx)),
        "max_lsb_diff": int(d.max()) if npx else 0,
        "offenders": [(int(y), int(x), a[y, x].tolist(), b[y, x].tolist()) for y, x in off[:16]],
        "max_colour": mx,
    }
