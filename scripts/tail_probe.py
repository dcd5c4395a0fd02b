"""Development aid: kernel time of 1/k of the bench frame (strip shards) -> fixed cost of the tail."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as graft
pkg = graft.load_package()
sph, lgt = pkg.synth_scene(1024, 4)
W, H, alias, S = 7680, 4320, 2.0, 8
strip = int(sys.argv[1]) if len(sys.argv) > 1 else 16
with pkg.Renderer(0) as r:
    r.upload_scene(sph, lgt)
    import os
    for kv in os.environ.get("RTG_OPTS", "").split(","):
        if "=" in kv:
            r.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    for k in (1, 8, 64):
        best = None
        for _ in range(2):
            r.render_strips(W, H, -4.0, alias, S, strip, 1 % k, k)
            st = r.stats()
            if best is None or st["kernel_ms"] < best["kernel_ms"]:
                best = st
        print(json.dumps({"k": k, "strip": strip, "kernel_ms": round(best["kernel_ms"], 3), "rays": best["rays"],
                          "ns_per_ray": round(best["kernel_ms"] * 1e6 / best["rays"], 4),
                          "fill": round(best["active_lane_iters"] / best["lane_iters"], 4), "rows": best["local_rows"],
                          "mean_tail_ms": round(best["phase_cycles"][0] / (best["grid"] * 8) / 1.965e6, 3),
                          "max_tail_ms": round(best["phase_cycles"][5] / 1.965e6, 3)}), flush=True)
