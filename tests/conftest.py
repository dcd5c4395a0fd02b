import ctypes
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run through gpurun)")


@pytest.fixture(scope="session")
def pkg():
    p = graft.load_package()
    p.build()
    return p


@pytest.fixture(scope="session")
def orc_mod():
    m = graft.load_oracle()
    m.build(reference=True)
    return m


@pytest.fixture(scope="session")
def oracle(orc_mod):
    return orc_mod.Oracle("port")


@pytest.fixture(scope="session")
def reference(orc_mod):
    """The reference's own headers compiled here (oracle/_ref); absent => skip."""
    if not orc_mod.reference_available(6):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    return orc_mod.Oracle("reference")


_HS_ARGS = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint,
            ctypes.c_uint, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_uint,
            ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]


@pytest.fixture(scope="session")
def hostsim():
    """CPU lane simulator of the kernel's state machine (tests/hostsim.cpp)."""
    lib = ctypes.CDLL(str(graft.build_hostsim()))
    lib.hostsim_render.argtypes = _HS_ARGS
    lib.hostsim_render.restype = ctypes.c_int

    def render(spheres, lights, width, height, zoom=-4.0, alias=1.0, max_stack=6, rows=None,
               no_filter=False, mode=None):
        begin, count, step = (0, height, 1) if rows is None else rows
        out = np.zeros((count, width, 3), np.float32)
        ctr = (ctypes.c_uint64 * 10)()
        spheres = np.ascontiguousarray(spheres)
        lights = np.ascontiguousarray(lights)
        rc = lib.hostsim_render(spheres.ctypes.data if len(spheres) else None, len(spheres),
                                lights.ctypes.data if len(lights) else None, len(lights),
                                width, height, zoom, alias, max_stack, begin, count, step,
                                out.ctypes.data, ctypes.addressof(ctr), int(no_filter) if mode is None else int(mode))
        assert rc == 0
        names = ["rays", "shadow_rays", "contain_queries", "contain_tests", "exact_tests", "samples",
                 "lane_iters", "active_lane_iters", "accel_violations", "cluster_tests"]
        return out, dict(zip(names, [int(v) for v in ctr]))

    return render
