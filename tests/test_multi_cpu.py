"""CPU test of the N>1 path's host logic with world_size 2 over gloo: strip partition,
max all-reduce on float bits, RGB8 all-gather and de-interleave.  The renderer stand-in
is the CPU lane simulator (test infrastructure); the GPU version of the same exchange is
exercised by bench.py --gpus N --verify."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent("""
    import os, sys, ctypes, importlib
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    import __graft_entry__ as graft
    pkg = graft.load_package(); om = graft.load_oracle()
    par = importlib.import_module(pkg.__name__ + ".parallel")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    W, H, alias, S, strip = 70, int(os.environ.get("RTG_TEST_H", "45")), 2.0, 6, 4
    sph, lgt = pkg.synth_scene(24, 3, seed=5)
    oracle = om.Oracle("port")
    rows = par.shard_rows(H, strip, rank, world)
    # render this rank's rows with the lane simulator, row by row
    lib = ctypes.CDLL(str(graft.build_hostsim()))
    part = np.zeros((len(rows), W, 3), np.float32)
    for i, rrow in enumerate(rows):
        one = np.zeros((1, W, 3), np.float32)
        rc = lib.hostsim_render(ctypes.c_void_p(sph.ctypes.data), len(sph), ctypes.c_void_p(lgt.ctypes.data), len(lgt),
                                W, H, ctypes.c_float(-4.0), ctypes.c_float(alias), S, int(rrow), 1, 1,
                                ctypes.c_void_p(one.ctypes.data), None, 0)
        assert rc == 0
        part[i] = one[0]
    local_max = np.float32(0)
    vals = part[~np.isnan(part)]
    if vals.size and vals.max() > 0: local_max = np.float32(vals.max())
    xchg = par.StripExchange(dist, torch, H, W, strip, rank, world, "cpu")
    bits = torch.tensor([int(np.float32(local_max).view(np.int32))], dtype=torch.int32)
    xchg.reduce_max(bits)
    gmax = np.int32(bits.item()).view(np.float32)
    gmax = np.float32(1.0) if gmax == 0 else gmax
    rgb = oracle.quantise(part, float(gmax))
    gathered = xchg.gather(torch.from_numpy(rgb.reshape(-1))).numpy()
    frame = par.assemble_host(gathered, H, W, strip, world, xchg.pitch)
    if rank == 0:
        full, _ = oracle.render(sph, lgt, W, H, -4.0, alias, S)
        mx = oracle.max_colour(full)
        assert np.float32(mx) == gmax, (mx, gmax)
        assert np.array_equal(frame, oracle.quantise(full, mx))
        print("MULTI_OK")
    dist.barrier(); dist.destroy_process_group()
""") % str(ROOT)


import pytest


@pytest.mark.parametrize("world,height", [(2, 45), (3, 50)])      # 50 rows / 4-row strips over 3 ranks: 20, 16 and 14 rows
def test_strip_exchange_gloo(tmp_path, pkg, orc_mod, world, height):
    graft = sys.modules["__graft_entry__"]
    graft.build_hostsim()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1", RTG_TEST_H=str(height))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(29533 + world), str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "MULTI_OK" in out.stdout


def test_assemble_mapping_matches_partition(pkg):
    import importlib
    par = importlib.import_module(pkg.__name__ + ".parallel")
    for H, strip, G in [(45, 4, 2), (101, 16, 8), (7, 3, 4), (64, 16, 4)]:
        for g in range(G):
            rows = par.shard_rows(H, strip, g, G)
            assert np.array_equal(rows, pkg.local_rows(H, strip, g, G))
            for lr, row in enumerate(rows):
                assert par.local_row_of(int(row), strip, G) == (g, lr)
        lay = par.shard_layout(H, 10, strip, G)
        assert sum(lay["rows"]) == H and lay["pitch"] % 16 == 0 and lay["pitch"] >= lay["max_rows"] * 30
