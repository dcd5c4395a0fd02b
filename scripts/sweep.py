"""BASELINE config 5: sphere-count sweep 16..4096 at 3840x2160, 1 spp, depth 8 — shared-memory (TMA bulk)
vs __constant__ staging.  Prints one JSON line per (N, staging)."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as graft

pkg = graft.load_package()
W, H, alias, S = 3840, 2160, 1.0, 8
r = pkg.Renderer(0)
peak = 148 * 128 * 2 * 1.965e9 / 1e12
print(json.dumps({"device": pkg.device_info(0), "ffma_peak_tflops": round(r.ffma_peak(8192), 2), "nominal_peak_tflops": round(peak, 2),
                  "workload": f"synth(N, 4 lights) {W}x{H} alias {alias:g} stack {S}"}))
for n in (16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
    sph, lgt = pkg.synth_scene(n, 4)
    r.upload_scene(sph, lgt)
    for staging in (1, 2, 3):                 # 3 = shared staging + the optional accelerated mode
        if staging == 1 and n > 1024:
            continue
        if staging == 3 and n < 64:
            continue
        r.set_option("staging", 2 if staging == 3 else staging)
        r.set_option("accel", 2 if staging == 3 else 0)
        best = None
        for _ in range(3):
            r.render(W, H, -4.0, alias, S)
            st = r.stats()
            if best is None or st["kernel_ms"] < best["kernel_ms"]:
                best = st
        r.set_option("staging", 0)
        r.set_option("accel", 0)
        ms = best["kernel_ms"]
        tests = (best["rays"] - best["null_rays"]) * n
        flops = 17.0 * tests + 8.0 * best["contain_tests"]
        print(json.dumps({"spheres": n, "staging": {1: "__constant__", 2: "shared (TMA bulk)", 3: "shared (TMA bulk) + accel (cluster filter; not brute force, frac is the contract's flops over time)"}[staging], "kernel_ms": round(ms, 3),
                          "frames_per_s": round(1e3 / ms, 2), "Mrays_per_s": round(best["rays"] / ms / 1e3, 1),
                          "tflops_17": round(flops / ms / 1e9, 2), "frac_of_nominal_peak": round(flops / ms / 1e9 / peak, 3),
                          "rays": best["rays"], "lane_utilisation": round(best["active_lane_iters"] / max(1, best["lane_iters"]), 3)}))
