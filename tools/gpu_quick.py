"""Quick GPU sanity + timing (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))   # golden_fixtures
import torch
from path_tracer_python_b200.engine import Engine, algorithmic_bytes
from path_tracer_python_b200 import scenes
import golden_fixtures

W = int(os.environ.get("W", 1920)); SPP = int(os.environ.get("SPP", 16))
e = Engine(0)
for k in ("smem_nodes", "block_threads"):
    if k.upper() in os.environ:
        e.set_option(k, int(os.environ[k.upper()]))
cs, meta = golden_fixtures.load_config("vol2_final_scene", W, 16 / 9)
info = e.set_scene(cs)
print("inner nodes", info.num_inner, "materials", info.num_materials, "depth", info.max_depth_tree, "bytes", info.total_bytes)
e.set_camera(meta["camera"], (0, 0, 0), 50)
H = meta["height"]
for mode in ("megakernel", "wavefront"):
    fn = e.render_megakernel if mode == "megakernel" else e.render_wavefront
    acc = e.new_accum(W, H)
    fn(acc, 0, 1, 7); torch.cuda.synchronize()
    acc.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(acc, 0, SPP, 7); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"{mode}: {W}x{H} spp={SPP}: {ms:.1f} ms  {W*H*SPP/ms/1e3:.1f} Mpaths/s  mean={acc.mean().item()/SPP:.5f}")
    if mode == "megakernel":
        mega = acc.clone()
e.enable_counters(True)
acc = e.new_accum(W, H); e.render_megakernel(acc, 0, 2, 7); c = e.counters()
print("counters (2 spp):", c, "alg bytes/path", algorithmic_bytes(c) / c["paths"], "rays/path", c["rays"] / c["paths"], "boxtests/ray", c["box_tests"]/c["rays"])
