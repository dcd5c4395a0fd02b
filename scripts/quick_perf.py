"""Quick perf probe (development aid): render a few workloads on cuda:0 and print kernel stats."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as graft

pkg = graft.load_package()
print(pkg.device_info(0))
r = pkg.Renderer(0)
print("ffma_peak TFLOP/s:", [round(r.ffma_peak(8192), 2) for _ in range(3)])
cases = [
    ("default 1080p a1 s4", pkg.default_scene(), 1920, 1080, 1.0, 4, {}),
    ("synth256 4K a1 s6", pkg.synth_scene(256, 4), 3840, 2160, 1.0, 6, {}),
    ("synth512 4K a1 s8", pkg.synth_scene(512, 4), 3840, 2160, 1.0, 8, {}),
    ("synth768 4K a1 s8", pkg.synth_scene(768, 4), 3840, 2160, 1.0, 8, {}),
    ("synth1024 4K a1 s8", pkg.synth_scene(1024, 4), 3840, 2160, 1.0, 8, {}),
    ("synth1024 4K a2 s8", pkg.synth_scene(1024, 4), 3840, 2160, 2.0, 8, {}),
    ("synth1024 8K a2 s8", pkg.synth_scene(1024, 4), 7680, 4320, 2.0, 8, {}),
    ("synth4096 2K a1 s8", pkg.synth_scene(4096, 4), 1920, 1080, 1.0, 8, {}),
    ("accel synth256 4K a1 s6", pkg.synth_scene(256, 4), 3840, 2160, 1.0, 6, {"accel": 2}),
    ("accel synth1024 4K a1 s8", pkg.synth_scene(1024, 4), 3840, 2160, 1.0, 8, {"accel": 1}),
    ("accel synth4096 2K a1 s8", pkg.synth_scene(4096, 4), 1920, 1080, 1.0, 8, {"accel": 1}),
]
if len(sys.argv) > 1:
    cases = [c for c in cases if any(a in c[0] for a in sys.argv[1:])]
import os
GLOBAL_OPTS = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in os.environ.get("RTG_OPTS", "").split(",") if "=" in kv}
for name, (sph, lgt), W, H, alias, S, opts in cases:
    opts = {**GLOBAL_OPTS, **opts}
    for k, v in opts.items():
        r.set_option(k, v)
    r.upload_scene(sph, lgt)
    best = None
    for it in range(3):
        r.render(W, H, -4.0, alias, S)
        st = r.stats()
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    for k in opts:
        r.set_option(k, 0)
    st = best
    ms = st["kernel_ms"]
    n = st["sph_num"]
    tests = (st["rays"] - st["null_rays"]) * n
    flops = 17.0 * tests + 8.0 * st["contain_tests"]
    print(json.dumps({
        "case": name, "ms": round(ms, 3), "Mrays/s": round(st["rays"] / ms / 1e3, 1),
        "Gtests/s": round(tests / ms / 1e6, 1), "TFLOP/s(17)": round(flops / ms / 1e9, 2),
        "frac_74.4": round(flops / ms / 1e9 / 74.4, 3),
        "fill": round(st["active_lane_iters"] / max(1, st["lane_iters"]), 3),
        "served_T/S/C": [st["served_trace"], st["served_shadow"], st["served_contain"]],
        "passes_T/S2/S4/C": [st["passes_trace"], st["passes_shadow2"], st["passes_shadow4"], st["passes_contain"]],
        "exact_per_query": round(st["exact_tests"] / max(1, st["active_lane_iters"]), 3),
        "phase_pct(refill+vote,setup,loop,resolve,advance)": [round(100.0 * c / max(1, sum(st["phase_cycles"])), 1) for c in st["phase_cycles"][:5]],
        "rays": st["rays"], "grid": st["grid"], "smem": st["smem_bytes"], "staging": st["staging"]}))
