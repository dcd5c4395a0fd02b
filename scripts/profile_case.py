"""One workload, a few frames — the command ncu wraps (scripts/profile_case.py N W H alias S [frames])."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as graft

n, W, H, alias, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
frames = int(sys.argv[6]) if len(sys.argv) > 6 else 2
pkg = graft.load_package()
sph, lgt = pkg.default_scene() if n == 0 else pkg.synth_scene(n, 4)
import os
with pkg.Renderer(0) as r:
    for kv in os.environ.get("RTG_OPTS", "").split(","):      # e.g. RTG_OPTS=slots=3,min_blocks=3
        if "=" in kv:
            k, v = kv.split("=")
            r.set_option(k, int(v))
    r.upload_scene(sph, lgt)
    for _ in range(frames):
        r.render(W, H, -4.0, alias, S)
        r.quantise(0.0)
        st = r.stats()
    print("kernel_ms", st["kernel_ms"], "rays", st["rays"], "grid", st["grid"], "smem", st["smem_bytes"])
