"""raytracer-gamma trace loop, B200-native — thin Python binding over the C-ABI.

The product is ``librt_cuda.so`` (``csrc/rt_shim.cu`` + ``csrc/rt_kernels.cuh``,
declared in ``include/rt_cuda.h``); this module only loads it with ctypes so that
tests, ``bench.py`` and the multi-GPU driver can call it.  There is no CPU
rendering path here: if the library is missing or no B200 is present the calls
raise.

The directory name contains a hyphen, so import it through
``__graft_entry__.load_package()`` (registers it as ``raytracer_gamma_b200``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
# Development overrides: RTG_LIB_DIR = a frozen copy of the built libraries + rt_gamma (scripts/freeze_build.sh;
# build() then never recompiles), RTG_LIB = another librt_cuda.so only.
_FROZEN = "RTG_LIB_DIR" in os.environ
_BIN_DIR = Path(os.environ["RTG_LIB_DIR"]).resolve() if _FROZEN else PKG_DIR
LIB_PATH = Path(os.environ.get("RTG_LIB", _BIN_DIR / "librt_cuda.so"))
SCENE_LIB_PATH = _BIN_DIR / "librt_scene.so"         # host-only: scene builders (no CUDA in it)
MULTI_LIB_PATH = _BIN_DIR / "librt_cuda_multi.so"    # multi-GPU sequencing over librt_cuda.so + NCCL
HOST_BIN = _BIN_DIR / "rt_gamma"

# Layout of the reference PODs (sphere.h:9-14, raytracer.h:20-25, vec.h:27-29)
SPHERE_DTYPE = np.dtype(
    [("pos", "<f4", 3), ("radius", "<f4"), ("matte", "<f4", 3), ("gloss", "<f4", 3),
     ("opacity", "<f4"), ("refractiveIndex", "<f4")])
LIGHT_DTYPE = np.dtype([("pos", "<f4", 3), ("col", "<f4", 3)])
assert SPHERE_DTYPE.itemsize == 48 and LIGHT_DTYPE.itemsize == 24

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off", "-std=c++17",
]


def _stale(out: Path, deps) -> bool:
    return not out.exists() or out.stat().st_mtime < max(p.stat().st_mtime for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile in-tree: librt_scene.so (gcc, host only), librt_cuda.so (nvcc, sm_100a),
    librt_cuda_multi.so (NCCL sequencing) and the C++ host program rt_gamma."""
    if _FROZEN:
        return LIB_PATH

    def run(cmd):
        if verbose:
            print(" ".join(map(str, cmd)))
        subprocess.run(list(map(str, cmd)), check=True)

    inc = list((REPO_ROOT / "include").glob("*.h"))
    scene_src = PKG_DIR / "host" / "rt_scene.c"
    if force or _stale(SCENE_LIB_PATH, [scene_src] + inc):
        run(["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fPIC", "-shared", f"-I{REPO_ROOT / 'include'}",
             "-o", SCENE_LIB_PATH, scene_src, "-lm"])
    lib = PKG_DIR / "librt_cuda.so"
    deps = [PKG_DIR / "csrc" / "rt_shim.cu"] + list((PKG_DIR / "csrc").glob("*.cuh")) + \
        list((PKG_DIR / "csrc").glob("*.h")) + inc
    if force or _stale(lib, deps):
        run(["nvcc", *NVCC_FLAGS, "-shared", f"-I{REPO_ROOT / 'include'}", f"-I{PKG_DIR / 'csrc'}",
             "-o", lib, PKG_DIR / "csrc" / "rt_shim.cu"])
    multi_src = PKG_DIR / "csrc" / "rt_multi.cpp"
    if force or _stale(MULTI_LIB_PATH, [multi_src, lib] + inc):
        run(["nvcc", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", f"-I{REPO_ROOT / 'include'}",
             "-o", MULTI_LIB_PATH, multi_src, f"-L{PKG_DIR}", "-lrt_cuda", "-lnccl",
             "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"])
    host_src = PKG_DIR / "host" / "main.cpp"
    host_deps = [host_src, PKG_DIR / "host" / "rt_png.h", lib, SCENE_LIB_PATH, MULTI_LIB_PATH] + inc
    if host_src.exists() and (force or _stale(HOST_BIN, host_deps)):
        run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", f"-I{REPO_ROOT / 'include'}",
             "-o", HOST_BIN, host_src, f"-L{PKG_DIR}", "-lrt_cuda_multi", "-lrt_cuda", "-lrt_scene",
             "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath-link,{PKG_DIR}"])
    return LIB_PATH


class RtCudaError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        name = _lib().rt_cuda_strerror(status).decode()
        super().__init__(f"{where}: {name} ({status}) {detail}".strip())


class Stats(ctypes.Structure):
    _fields_ = [
        ("rays", ctypes.c_uint64), ("shadow_rays", ctypes.c_uint64),
        ("contain_queries", ctypes.c_uint64), ("contain_tests", ctypes.c_uint64),
        ("exact_tests", ctypes.c_uint64), ("samples", ctypes.c_uint64),
        ("lane_iters", ctypes.c_uint64), ("active_lane_iters", ctypes.c_uint64),
        ("served_trace", ctypes.c_uint64), ("served_shadow", ctypes.c_uint64),
        ("served_contain", ctypes.c_uint64), ("passes", ctypes.c_uint64),
        ("passes_trace", ctypes.c_uint64), ("passes_shadow2", ctypes.c_uint64),
        ("passes_shadow4", ctypes.c_uint64), ("passes_contain", ctypes.c_uint64),
        ("phase_cycles", ctypes.c_uint64 * 6),
        ("filter_tests", ctypes.c_uint64), ("null_rays", ctypes.c_uint64),
        ("sph_num", ctypes.c_uint32), ("sph_padded", ctypes.c_uint32), ("lgt_num", ctypes.c_uint32),
        ("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("local_rows", ctypes.c_uint32),
        ("kernel_ms", ctypes.c_float), ("max_colour", ctypes.c_float),
        ("kernel_launches", ctypes.c_uint32),
        ("grid", ctypes.c_uint32), ("block", ctypes.c_uint32), ("smem_bytes", ctypes.c_uint32),
        ("staging", ctypes.c_uint32), ("engine", ctypes.c_uint32),
        ("accel", ctypes.c_uint32), ("clusters", ctypes.c_uint32), ("slots_on_chip", ctypes.c_uint32),
    ]

    def as_dict(self) -> dict:
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["phase_cycles"] = list(self.phase_cycles)
        return d


# Every symbol include/rt_cuda.h and include/rt_scene.h declare: (name, restype, argtypes)
_P = ctypes.c_void_p
C_ABI = [
    ("rt_cuda_init", ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_P)]),
    ("rt_cuda_upload_scene", ctypes.c_int, [_P, _P, ctypes.c_uint, _P, ctypes.c_uint]),
    ("rt_cuda_render", ctypes.c_int,
     [_P, ctypes.c_uint, ctypes.c_uint, ctypes.c_float, ctypes.c_float, ctypes.c_int]),
    ("rt_cuda_render_strips", ctypes.c_int,
     [_P, ctypes.c_uint, ctypes.c_uint, ctypes.c_float, ctypes.c_float, ctypes.c_int,
      ctypes.c_uint, ctypes.c_uint, ctypes.c_uint]),
    ("rt_cuda_readback", ctypes.c_int, [_P, _P, ctypes.POINTER(ctypes.c_float)]),
    ("rt_cuda_readback_rgb8", ctypes.c_int, [_P, _P, ctypes.c_float]),
    ("rt_cuda_quantise", ctypes.c_int, [_P, ctypes.c_float]),
    ("rt_cuda_quantise_to", ctypes.c_int, [_P, _P, ctypes.c_size_t, ctypes.c_float]),
    ("rt_cuda_readback_rgb8_async", ctypes.c_int, [_P, _P, ctypes.c_float, ctypes.POINTER(ctypes.c_int)]),
    ("rt_cuda_readback_wait", ctypes.c_int, [_P, ctypes.c_int]),
    ("rt_cuda_host_alloc", _P, [ctypes.c_size_t]),
    ("rt_cuda_host_free", None, [_P]),
    ("rt_cuda_get_stream", _P, [_P]),
    ("rt_cuda_get_device", ctypes.c_int, [_P]),
    ("rt_cuda_flush_l2", ctypes.c_int, [_P]),
    ("rt_cuda_device_packed", _P, [_P]),
    ("rt_cuda_device_rgb8", _P, [_P]),
    ("rt_cuda_device_max", _P, [_P]),
    ("rt_cuda_pack", ctypes.c_int, [_P]),
    ("rt_cuda_set_stream", ctypes.c_int, [_P, _P]),
    ("rt_cuda_synchronize", ctypes.c_int, [_P]),
    ("rt_cuda_set_option", ctypes.c_int, [_P, ctypes.c_char_p, ctypes.c_long]),
    ("rt_cuda_get_stats", ctypes.c_int, [_P, ctypes.POINTER(Stats)]),
    ("rt_cuda_assemble_rgb8", ctypes.c_int,
     [_P, _P, _P, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_size_t]),
    ("rt_cuda_ffma_peak", ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]),
    ("rt_cuda_destroy", None, [_P]),
    ("rt_cuda_strerror", ctypes.c_char_p, [ctypes.c_int]),
    ("rt_cuda_last_error", ctypes.c_char_p, [_P]),
    ("rt_cuda_device_count", ctypes.c_int, []),
    ("rt_cuda_device_info", ctypes.c_int, [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]),
]

# include/rt_scene.h (librt_scene.so, host only)
SCENE_ABI = [
    ("rt_make_material", None,
     [_P, _P, _P, ctypes.c_float, ctypes.c_float, ctypes.c_float]),
    ("rt_scene_default", None, [_P, _P]),
    ("rt_scene_synth", ctypes.c_int, [ctypes.c_uint, ctypes.c_uint, ctypes.c_uint64, _P, _P]),
    ("rt_scene_save", ctypes.c_int, [ctypes.c_char_p, _P, ctypes.c_uint, _P, ctypes.c_uint]),
    ("rt_scene_load", ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_uint),
                                     ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_uint)]),
    ("rt_scene_free", None, [_P]),
]

# include/rt_cuda_multi.h (librt_cuda_multi.so)
MULTI_ABI = [
    ("rt_cuda_multi_shard_rows", ctypes.c_uint, [ctypes.c_uint] * 4),
    ("rt_cuda_multi_shard_pitch", ctypes.c_size_t, [ctypes.c_uint] * 4),
    ("rt_cuda_multi_locate_row", None, [ctypes.c_uint] * 3 + [ctypes.POINTER(ctypes.c_uint)] * 2),
    ("rt_cuda_multi_init", ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(_P)]),
    ("rt_cuda_multi_unique_id", ctypes.c_int, [_P, ctypes.c_size_t]),
    ("rt_cuda_multi_init_rank", ctypes.c_int, [ctypes.c_int, _P, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                               ctypes.POINTER(_P)]),
    ("rt_cuda_multi_world_size", ctypes.c_int, [_P]),
    ("rt_cuda_multi_local_count", ctypes.c_int, [_P]),
    ("rt_cuda_multi_context", _P, [_P, ctypes.c_int]),
    ("rt_cuda_multi_upload_scene", ctypes.c_int, [_P, _P, ctypes.c_uint, _P, ctypes.c_uint]),
    ("rt_cuda_multi_set_option", ctypes.c_int, [_P, ctypes.c_char_p, ctypes.c_long]),
    ("rt_cuda_multi_render", ctypes.c_int,
     [_P, ctypes.c_uint, ctypes.c_uint, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_uint]),
    ("rt_cuda_multi_synchronize", ctypes.c_int, [_P]),
    ("rt_cuda_multi_device_frame", _P, [_P, ctypes.c_int]),
    ("rt_cuda_multi_readback_rgb8", ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.POINTER(ctypes.c_float)]),
    ("rt_cuda_multi_readback_rgb8_async", ctypes.c_int, [_P, ctypes.c_int, _P]),
    ("rt_cuda_multi_readback_wait", ctypes.c_int, [_P, ctypes.c_int]),
    ("rt_cuda_multi_step_ms", ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]),
    ("rt_cuda_multi_flush_l2", ctypes.c_int, [_P]),
    ("rt_cuda_multi_last_error", ctypes.c_char_p, [_P]),
    ("rt_cuda_multi_destroy", None, [_P]),
]

_LIB = None
_SCENE_LIB = None
_MULTI_LIB = None


def _bind(path: Path, table) -> ctypes.CDLL:
    if not path.exists():
        raise FileNotFoundError(
            f"{path} is missing: run __graft_entry__.build() (nvcc, sm_100a). "
            "There is no CPU fallback for the trace loop.")
    lib = ctypes.CDLL(str(path), mode=ctypes.RTLD_GLOBAL)
    for name, res, args in table:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def _lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = _bind(LIB_PATH, C_ABI)
    return _LIB


def _scene_lib() -> ctypes.CDLL:
    global _SCENE_LIB
    if _SCENE_LIB is None:
        _SCENE_LIB = _bind(SCENE_LIB_PATH, SCENE_ABI)
    return _SCENE_LIB


def _multi_lib() -> ctypes.CDLL:
    """librt_cuda_multi.so: needs librt_cuda.so (loaded first, so an RTG_LIB override is honoured) and NCCL."""
    global _MULTI_LIB
    if _MULTI_LIB is None:
        _lib()
        _MULTI_LIB = _bind(MULTI_LIB_PATH, MULTI_ABI)
    return _MULTI_LIB


def load() -> ctypes.CDLL:
    return _lib()


def device_count() -> int:
    return _lib().rt_cuda_device_count()


def device_info(device: int = 0) -> str:
    buf = ctypes.create_string_buffer(512)
    rc = _lib().rt_cuda_device_info(device, buf, len(buf))
    if rc:
        raise RtCudaError(rc, "rt_cuda_device_info")
    return buf.value.decode()


# ---- scenes (host side, include/rt_scene.h) ---------------------------------
def default_scene():
    """The scene literal of main.cpp:113-168 as (spheres, lights) structured arrays."""
    sph = np.zeros(3, SPHERE_DTYPE)
    lgt = np.zeros(2, LIGHT_DTYPE)
    _scene_lib().rt_scene_default(sph.ctypes.data, lgt.ctypes.data)
    return sph, lgt


def synth_scene(n: int, lights: int = 4, seed: int = 0):
    """synth(N, L, seed) of SURVEY.md §8(d)."""
    sph = np.zeros(n, SPHERE_DTYPE)
    lgt = np.zeros(lights, LIGHT_DTYPE)
    rc = _scene_lib().rt_scene_synth(n, lights, seed, sph.ctypes.data, lgt.ctypes.data)
    if rc:
        raise ValueError("rt_scene_synth: bad arguments")
    return sph, lgt


def save_scene(path, spheres: np.ndarray, lights: np.ndarray) -> None:
    spheres = np.ascontiguousarray(spheres)
    lights = np.ascontiguousarray(lights)
    rc = _scene_lib().rt_scene_save(str(path).encode(), spheres.ctypes.data if len(spheres) else None, len(spheres),
                              lights.ctypes.data if len(lights) else None, len(lights))
    if rc:
        raise OSError(f"rt_scene_save({path}) failed")


def load_scene(path):
    ps, pl = _P(), _P()
    ns, nl = ctypes.c_uint(0), ctypes.c_uint(0)
    rc = _scene_lib().rt_scene_load(str(path).encode(), ctypes.byref(ps), ctypes.byref(ns), ctypes.byref(pl), ctypes.byref(nl))
    if rc:
        raise OSError(f"rt_scene_load({path}) failed")
    try:
        sph = np.zeros(ns.value, SPHERE_DTYPE)
        lgt = np.zeros(nl.value, LIGHT_DTYPE)
        if ns.value:
            ctypes.memmove(sph.ctypes.data, ps, sph.nbytes)
        if nl.value:
            ctypes.memmove(lgt.ctypes.data, pl, lgt.nbytes)
    finally:
        _scene_lib().rt_scene_free(ps)
        _scene_lib().rt_scene_free(pl)
    return sph, lgt


def make_material(matte, gloss, opacity, gloss_factor, refractive_index) -> np.ndarray:
    out = np.zeros(8, np.float32)
    m = np.asarray(matte, np.float32)
    g = np.asarray(gloss, np.float32)
    _scene_lib().rt_make_material(out.ctypes.data, m.ctypes.data, g.ctypes.data, opacity, gloss_factor,
                            refractive_index)
    return out


def local_rows(height: int, strip_rows: int, first: int, stride: int) -> np.ndarray:
    """Global row numbers, in storage order, of the shard (first, stride) of a strip-interleaved frame."""
    rows = np.arange(height)
    return rows[(rows // strip_rows) % stride == first]


class Renderer:
    """One context of the C-ABI (one GPU).  Mirrors the call sequence of the
    reference's main(): init -> upload_scene -> render -> readback."""

    def __init__(self, device: int = 0):
        self._lib = _lib()
        self._ctx = _P()
        rc = self._lib.rt_cuda_init(device, ctypes.byref(self._ctx))
        if rc:
            raise RtCudaError(rc, "rt_cuda_init")
        self.device = device
        self.width = self.height = self.rows = 0

    def _check(self, rc: int, where: str):
        if rc:
            raise RtCudaError(rc, where, self._lib.rt_cuda_last_error(self._ctx).decode())

    def close(self):
        if self._ctx:
            self._lib.rt_cuda_destroy(self._ctx)
            self._ctx = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, key: str, value: int):
        self._check(self._lib.rt_cuda_set_option(self._ctx, key.encode(), int(value)), "rt_cuda_set_option")

    def set_stream(self, cuda_stream: int):
        self._check(self._lib.rt_cuda_set_stream(self._ctx, _P(cuda_stream)), "rt_cuda_set_stream")

    def upload_scene(self, spheres: np.ndarray, lights: np.ndarray):
        spheres = np.ascontiguousarray(spheres)
        lights = np.ascontiguousarray(lights)
        assert spheres.dtype.itemsize == 48 or spheres.size == 0
        assert lights.dtype.itemsize == 24 or lights.size == 0
        self._check(self._lib.rt_cuda_upload_scene(
            self._ctx, spheres.ctypes.data if len(spheres) else None, len(spheres),
            lights.ctypes.data if len(lights) else None, len(lights)), "rt_cuda_upload_scene")

    def render(self, width: int, height: int, zoom: float = -4.0, alias: float = 1.0, max_stack: int = 6):
        self._check(self._lib.rt_cuda_render(self._ctx, width, height, zoom, alias, max_stack), "rt_cuda_render")
        self.width, self.height, self.rows = width, height, height

    def render_strips(self, width, height, zoom, alias, max_stack, strip_rows, first, stride):
        self._check(self._lib.rt_cuda_render_strips(self._ctx, width, height, zoom, alias, max_stack,
                                                    strip_rows, first, stride), "rt_cuda_render_strips")
        self.width, self.height = width, height
        self.rows = len(local_rows(height, strip_rows, first, stride))

    def readback(self, out: np.ndarray | None = None):
        """-> (float32 [rows, W, 3], max colour)"""
        if out is None:
            out = np.empty((self.rows, self.width, 3), np.float32)
        mx = ctypes.c_float(0)
        self._check(self._lib.rt_cuda_readback(self._ctx, out.ctypes.data, ctypes.byref(mx)), "rt_cuda_readback")
        return out, mx.value

    def readback_rgb8(self, max_colour: float = 0.0, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.rows, self.width, 3), np.uint8)
        self._check(self._lib.rt_cuda_readback_rgb8(self._ctx, out.ctypes.data, max_colour), "rt_cuda_readback_rgb8")
        return out

    def quantise(self, max_colour: float = 0.0):
        self._check(self._lib.rt_cuda_quantise(self._ctx, max_colour), "rt_cuda_quantise")

    def quantise_to(self, dev_ptr: int, nbytes: int, max_colour: float = 0.0):
        self._check(self._lib.rt_cuda_quantise_to(self._ctx, _P(dev_ptr), nbytes, max_colour), "rt_cuda_quantise_to")

    def readback_rgb8_async(self, out: "HostBuffer | np.ndarray", max_colour: float = 0.0) -> int:
        """Enqueue quantise + D2H into `out` (a HostBuffer = pinned, or any uint8 array) and return a ticket."""
        arr = out.array if isinstance(out, HostBuffer) else out
        t = ctypes.c_int(-1)
        self._check(self._lib.rt_cuda_readback_rgb8_async(self._ctx, arr.ctypes.data, max_colour, ctypes.byref(t)),
                    "rt_cuda_readback_rgb8_async")
        return t.value

    def readback_wait(self, ticket: int):
        self._check(self._lib.rt_cuda_readback_wait(self._ctx, ticket), "rt_cuda_readback_wait")

    def flush_l2(self):
        self._check(self._lib.rt_cuda_flush_l2(self._ctx), "rt_cuda_flush_l2")

    def stream(self) -> int:
        return int(self._lib.rt_cuda_get_stream(self._ctx) or 0)

    def pack(self):
        self._check(self._lib.rt_cuda_pack(self._ctx), "rt_cuda_pack")

    def synchronize(self):
        self._check(self._lib.rt_cuda_synchronize(self._ctx), "rt_cuda_synchronize")

    def stats(self) -> dict:
        s = Stats()
        self._check(self._lib.rt_cuda_get_stats(self._ctx, ctypes.byref(s)), "rt_cuda_get_stats")
        return s.as_dict()

    def assemble_rgb8(self, gathered_ptr: int, out_ptr: int, width, height, strip_rows, shards, pitch):
        self._check(self._lib.rt_cuda_assemble_rgb8(self._ctx, _P(gathered_ptr), _P(out_ptr), width, height,
                                                    strip_rows, shards, pitch), "rt_cuda_assemble_rgb8")

    def ffma_peak(self, iters: int = 4096) -> float:
        t = ctypes.c_float(0)
        self._check(self._lib.rt_cuda_ffma_peak(self._ctx, iters, ctypes.byref(t)), "rt_cuda_ffma_peak")
        return t.value

    def device_ptr(self, which: str) -> int:
        fn = {"packed": self._lib.rt_cuda_device_packed, "rgb8": self._lib.rt_cuda_device_rgb8,
              "max": self._lib.rt_cuda_device_max}[which]
        return int(fn(self._ctx) or 0)


class HostBuffer:
    """Page-locked host memory from rt_cuda_host_alloc, viewed as a uint8 numpy array."""

    def __init__(self, nbytes: int):
        self._lib = _lib()
        self.ptr = self._lib.rt_cuda_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError(f"rt_cuda_host_alloc({nbytes}) failed")
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.rt_cuda_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def multi_layout(width: int, height: int, strip_rows: int, world: int) -> dict:
    """Rows per rank and the gather pitch as librt_cuda_multi.so computes them (host arithmetic only)."""
    lib = _multi_lib()
    return {"rows": [lib.rt_cuda_multi_shard_rows(height, strip_rows, g, world) for g in range(world)],
            "pitch": lib.rt_cuda_multi_shard_pitch(width, height, strip_rows, world)}


def multi_locate_row(row: int, strip_rows: int, world: int):
    r, lr = ctypes.c_uint(0), ctypes.c_uint(0)
    _multi_lib().rt_cuda_multi_locate_row(row, strip_rows, world, ctypes.byref(r), ctypes.byref(lr))
    return r.value, lr.value


def multi_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    rc = _multi_lib().rt_cuda_multi_unique_id(buf, 128)
    if rc:
        raise RtCudaError(rc, "rt_cuda_multi_unique_id")
    return buf.raw


class MultiRenderer:
    """One frame sharded by row strips over several GPUs (include/rt_cuda_multi.h).

    MultiRenderer(gpus=N)                      one process drives N devices (ncclCommInitAll)
    MultiRenderer(rank=r, world=G, uid=bytes, device=d)   one process per GPU (torchrun)"""

    def __init__(self, gpus: int | None = None, devices=None, *, rank: int | None = None, world: int | None = None,
                 uid: bytes | None = None, device: int = 0):
        self._lib = _multi_lib()
        self._m = _P()
        if rank is None:
            arr = (ctypes.c_int * gpus)(*devices) if devices is not None else None
            rc = self._lib.rt_cuda_multi_init(gpus, arr, ctypes.byref(self._m))
            where = "rt_cuda_multi_init"
        else:
            rc = self._lib.rt_cuda_multi_init_rank(device, uid, len(uid), rank, world, ctypes.byref(self._m))
            where = "rt_cuda_multi_init_rank"
        if rc:
            raise RtCudaError(rc, where)
        self.world = self._lib.rt_cuda_multi_world_size(self._m)
        self.local = self._lib.rt_cuda_multi_local_count(self._m)
        self.width = self.height = 0

    def _check(self, rc: int, where: str):
        if rc:
            raise RtCudaError(rc, where, self._lib.rt_cuda_multi_last_error(self._m).decode())

    def close(self):
        if self._m:
            self._lib.rt_cuda_multi_destroy(self._m)
            self._m = _P()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_scene(self, spheres: np.ndarray, lights: np.ndarray):
        spheres = np.ascontiguousarray(spheres)
        lights = np.ascontiguousarray(lights)
        self._check(self._lib.rt_cuda_multi_upload_scene(
            self._m, spheres.ctypes.data if len(spheres) else None, len(spheres),
            lights.ctypes.data if len(lights) else None, len(lights)), "rt_cuda_multi_upload_scene")

    def set_option(self, key: str, value: int):
        self._check(self._lib.rt_cuda_multi_set_option(self._m, key.encode(), int(value)), "rt_cuda_multi_set_option")

    def render(self, width, height, zoom=-4.0, alias=1.0, max_stack=6, strip_rows=0):
        self._check(self._lib.rt_cuda_multi_render(self._m, width, height, zoom, alias, max_stack, strip_rows),
                    "rt_cuda_multi_render")
        self.width, self.height = width, height

    def synchronize(self):
        self._check(self._lib.rt_cuda_multi_synchronize(self._m), "rt_cuda_multi_synchronize")

    def flush_l2(self):
        self._check(self._lib.rt_cuda_multi_flush_l2(self._m), "rt_cuda_multi_flush_l2")

    def readback_rgb8(self, local: int = 0, out: np.ndarray | None = None):
        """-> (uint8 [H, W, 3] assembled frame, global max colour)"""
        if out is None:
            out = np.empty((self.height, self.width, 3), np.uint8)
        mx = ctypes.c_float(0)
        self._check(self._lib.rt_cuda_multi_readback_rgb8(self._m, local, out.ctypes.data, ctypes.byref(mx)),
                    "rt_cuda_multi_readback_rgb8")
        return out, mx.value

    def readback_rgb8_async(self, out: "HostBuffer", local: int = 0):
        self._check(self._lib.rt_cuda_multi_readback_rgb8_async(self._m, local, out.array.ctypes.data),
                    "rt_cuda_multi_readback_rgb8_async")

    def readback_wait(self, local: int = 0):
        self._check(self._lib.rt_cuda_multi_readback_wait(self._m, local), "rt_cuda_multi_readback_wait")

    def set_stream(self, cuda_stream: int, local: int = 0):
        ctx = self._lib.rt_cuda_multi_context(self._m, local)
        rc = _lib().rt_cuda_set_stream(_P(ctx), _P(cuda_stream))
        if rc:
            raise RtCudaError(rc, "rt_cuda_set_stream")

    def step_ms(self, local: int = 0) -> float:
        ms = ctypes.c_float(0)
        self._check(self._lib.rt_cuda_multi_step_ms(self._m, local, ctypes.byref(ms)), "rt_cuda_multi_step_ms")
        return ms.value

    def device_frame(self, local: int = 0) -> int:
        return int(self._lib.rt_cuda_multi_device_frame(self._m, local) or 0)

    def stats(self, local: int = 0) -> dict:
        ctx = self._lib.rt_cuda_multi_context(self._m, local)
        s = Stats()
        rc = _lib().rt_cuda_get_stats(_P(ctx), ctypes.byref(s))
        if rc:
            raise RtCudaError(rc, "rt_cuda_get_stats")
        return s.as_dict()


def write_ppm(path, rgb8: np.ndarray):
    """Binary P6 with the reference's header (main.cpp:66)."""
    h, w, _ = rgb8.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(rgb8).tobytes())
