"""CPU tests of bench.py's reference arm (the one leg of the benchmark that runs without a GPU)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

PROBE = """
import json, os, sys
sys.argv = ["bench.py", "--impl", "reference", "--workload", "config2", "--width", "160", "--height", "90",
            "--steps", "2", "--warmup", "1", "--cpu-seconds", "0.2"]
sys.path.insert(0, %r)
import bench
rc = bench.main()
maps = open("/proc/self/maps").read()
print(json.dumps({"rc": rc, "maps_cuda": [l.split()[-1] for l in maps.splitlines() if "librt_cuda" in l or "libcudart" in l or "libnccl" in l],
                  "maps_scene": any("librt_scene" in l for l in maps.splitlines())}))
""" % str(ROOT)


def test_reference_arm_uses_all_cores_under_torchrun_env_and_maps_no_cuda_library(pkg, orc_mod):
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm must still use every core the
    process may run on (the per-N reference ran on one thread), must report what it used, and
    must not map the CUDA libraries (its scenes come from the host-only librt_scene.so)."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, "-c", PROBE], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    line, probe = json.loads(lines[0]), json.loads(lines[1])
    assert line["impl"] == "reference" and line["metric"] == "Mrays/s" and line["value"] > 0
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and len(line["seconds_per_step"]) == 2
    assert probe["rc"] == 0 and probe["maps_cuda"] == [] and probe["maps_scene"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
