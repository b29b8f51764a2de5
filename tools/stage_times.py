"""Per-stage device times of one wavefront batch on config 2 (pt_enable_stage_timing), no profiler.
    python tools/stage_times.py            env: SPP (default 32), W, PT_LIB_PATH, OPTS=k=v,..."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from path_tracer_python_b200.engine import Engine
from path_tracer_python_b200 import scenes
W = int(os.environ.get("W", 1920)); SPP = int(os.environ.get("SPP", 32))
e = Engine(0)
cs, meta = scenes.load_config(os.environ.get("SCENE", "vol2_final_scene"), W, 16 / 9)
e.set_scene(cs); e.set_camera(meta["camera"], (0, 0, 0), 50)
H = meta["height"]
for k, v in [kv.split("=") for kv in os.environ.get("OPTS", "").split(",") if kv]:
    e.set_option(k, int(v))
acc = e.new_accum(W, H)
e.render_wavefront(acc, 0, SPP, 7); torch.cuda.synchronize()
e.enable_stage_timing(True)
best = None
for _ in range(3):
    acc.zero_(); e.render_wavefront(acc, 0, SPP, 7)
    t = e.stage_times()
    if best is None or sum(x[2] for x in t) < sum(x[2] for x in best):
        best = t
tot = collections.OrderedDict()
for name, wave, ms in best:
    tot[name] = tot.get(name, 0.0) + ms
total = sum(tot.values())
print(f"{W}x{H} spp={SPP}: sum of stages {total:.2f} ms = {W*H*SPP/total/1e3:.0f} Mpaths/s")
print("  totals: " + "  ".join(f"{k} {v:.2f}" for k, v in tot.items()))
for stage in ("extend", "bin", "shade"):
    print(f"  {stage:7s} per wave: " + " ".join(f"{ms:.2f}" for n, w, ms in best if n == stage))
