import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))   # golden_fixtures
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
from path_tracer_python_b200.engine import Engine
from path_tracer_python_b200 import scenes
import golden_fixtures
import oracle as O
name = sys.argv[1] if len(sys.argv) > 1 else "wavefront_comparison"
e = Engine(0)
cs, meta, out = golden_fixtures.load_fixture(name)
e.set_scene(cs); e.set_camera(meta["camera"], meta["background"], meta["max_depth"])
W, H, spp, seed = meta["width"], meta["height"], meta["spp"], meta["rng_seed"]
S = O.OracleScene(cs); cam = O.make_camera(meta["camera"], meta["background"], meta["max_depth"])
for sort in (0, 1):
    e.set_option("wavefront_sort", sort)
    res = []
    for rep in range(2):
        per = []
        for s in range(spp):
            acc = e.new_accum(W, H); e.render_wavefront(acc, s, 1, seed); per.append(acc.cpu().numpy())
        res.append(per)
    print("sort", sort, "deterministic:", all(np.array_equal(a, b) for a, b in zip(*res)))
    for s in range(spp):
        ref, _ = O.render(S, cam, W, H, 1, seed=seed, mode="wavefront", sample_begin=s)
        bad = np.argwhere(np.abs(res[0][s] - ref).max(axis=2) > 0)
        print(" sample", s, "mismatch px:", [(int(y), int(x), res[0][s][y, x].tolist(), ref[y, x].tolist()) for y, x in bad][:6])
