"""Short, deterministic run for ncu: 2 warm-up launches + 1 measured launch of the megakernel, then
the wavefront pipeline, on config 2 (vol2_final_scene 1920x1080) at SPP samples (default 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))   # golden_fixtures
import torch
from path_tracer_python_b200.engine import Engine
from path_tracer_python_b200 import scenes
import golden_fixtures

W = int(os.environ.get("W", 1920)); SPP = int(os.environ.get("SPP", 4)); MODES = os.environ.get("MODES", "megakernel,wavefront").split(",")
e = Engine(0)
for k in ("smem_nodes", "block_threads"):
    if k.upper() in os.environ:
        e.set_option(k, int(os.environ[k.upper()]))
for k, v in [kv.split("=") for kv in os.environ.get("OPTS", "").split(",") if kv]:     # e.g. OPTS=wf_overlap=0: one pipeline, one batch
    e.set_option(k, int(v))
if os.environ.get("FLIGHT"):
    e.wavefront_samples_in_flight = int(os.environ["FLIGHT"])
cs, meta = golden_fixtures.load_config("vol2_final_scene", W, 16 / 9)
e.set_scene(cs); e.set_camera(meta["camera"], (0, 0, 0), 50)
H = meta["height"]
for mode in MODES:
    fn = e.render_megakernel if mode == "megakernel" else e.render_wavefront
    acc = e.new_accum(W, H)
    for _ in range(2):
        fn(acc, 0, SPP if mode == "megakernel" else 1, 7)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(acc, 0, SPP, 7); b.record(); torch.cuda.synchronize()
    print(f"{mode}: {W}x{H} spp={SPP}: {a.elapsed_time(b):.2f} ms  {W*H*SPP/a.elapsed_time(b)/1e3:.1f} Mpaths/s")
