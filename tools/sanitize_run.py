"""Tiny deterministic workload for compute-sanitizer (memcheck / racecheck): every kernel, both
architectures, compat and full feature mode, on the small golden scenes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))   # golden_fixtures
import torch
from path_tracer_python_b200.engine import Engine
from path_tracer_python_b200 import scenes
import golden_fixtures
e = Engine(0)
for name in ("feature_zoo", "cornell_smoke", "full_features"):
    cs, meta, out = golden_fixtures.load_fixture(name)
    e.set_scene(cs); e.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    W, H = meta["width"], meta["height"]
    o, d = e.generate_camera_rays(W, H, 1, 0)
    e.first_hit(o, d)
    for fn in (e.render_megakernel, e.render_wavefront):
        acc = e.new_accum(W, H); fn(acc, 0, 2, 3)
    img = e.tonemap(acc, 2)
    torch.cuda.synchronize()
    print(name, "ok", float(acc.mean()))
