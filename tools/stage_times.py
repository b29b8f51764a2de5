"""Per-stage device times of one wavefront batch on config 2 (pt_enable_stage_timing), no profiler.
    python tools/stage_times.py            env: SPP (default 32), W, FLIGHT (samples per batch), PT_LIB_PATH, OPTS=k=v,..."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))   # golden_fixtures
import torch
from path_tracer_python_b200.engine import Engine
from path_tracer_python_b200 import scenes
import golden_fixtures
W = int(os.environ.get("W", 1920)); SPP = int(os.environ.get("SPP", 32))
e = Engine(0)
if os.environ.get("SCENE") == "cornell4":        # BASELINE config 4: Cornell + medium + glass + SSS-as-Lambert, 1024^2
    from path_tracer_python_b200 import scene_compiler as SC, bvh
    W = int(os.environ.get("W", 1024))
    world, cam = scenes.cornell_config4(W, 64); cam.initialize()
    cs = SC.compile_scene(world); cs.update(bvh.compile_bvh(cs)); cs.update(SC.perlin_tables())
    v = lambda a: [a.x, a.y, a.z]
    meta = {"height": cam.img_height,
            "camera": {"cam_center": v(cam.center), "cam_pixel00": v(cam.pixel00_loc), "cam_delta_u": v(cam.delta_u), "cam_delta_v": v(cam.delta_v),
                       "cam_defocus_disk_u": v(cam.defocus_disk_u), "cam_defocus_disk_v": v(cam.defocus_disk_v), "cam_defocus_angle": 0.0}}
elif os.environ.get("SCENE") == "heightfield":  # BASELINE config 3: 1 M triangles
    cs, meta = scenes.heightfield_scene()
    W = meta["width"]
else:
    cs, meta = golden_fixtures.load_config(os.environ.get("SCENE", "vol2_final_scene"), W, 16 / 9)
e.set_scene(cs); e.set_camera(meta["camera"], meta.get("background", (0, 0, 0)), meta.get("max_depth", 50))
H = meta["height"]
for k, v in [kv.split("=") for kv in os.environ.get("OPTS", "").split(",") if kv]:
    e.set_option(k, int(v))
if os.environ.get("FLIGHT"):
    e.wavefront_samples_in_flight = int(os.environ["FLIGHT"])      # samples of every pixel per batch (default: auto)
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
for stage in ("extend", "bin", "shade", "tail"):
    print(f"  {stage:7s} per wave: " + " ".join(f"{ms:.2f}" for n, w, ms in best if n == stage))
