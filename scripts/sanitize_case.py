"""Development aid: a few small renders in every mode, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as graft
pkg = graft.load_package()
with pkg.Renderer(0) as r:
    for n, W, H, alias, S, opts in [(0, 96, 64, 2.0, 6, {}), (40, 80, 48, 1.0, 8, {}), (300, 64, 40, 2.0, 8, {}),
                                    (300, 64, 40, 1.0, 8, {"accel": 2}), (300, 48, 32, 1.0, 6, {"staging": 1}),
                                    (300, 48, 32, 1.0, 6, {"no_filter": 1}), (1100, 48, 32, 1.0, 8, {"accel": 1})]:
        sph, lgt = pkg.default_scene() if n == 0 else pkg.synth_scene(n, 4)
        for k, v in opts.items():
            r.set_option(k, v)
        r.upload_scene(sph, lgt)
        r.render(W, H, -4.0, alias, S)
        fb, mx = r.readback()
        r.render_strips(W, H, -4.0, alias, S, 4, 1, 3)
        rgb = r.readback_rgb8(0.0)
        st = r.stats()
        for k in opts:
            r.set_option(k, 0)
        print(n, W, H, opts, "rays", st["rays"], "max", mx, "engine", st["engine"], "accel", st["accel"])
print("done")
