"""One context renders a sequence of scenes (vol2 / hf / cornell ...), two calls each, with several
batches per call -- the stress sequence that exposed the device-allocation problem described in
DESIGN.md section 5.2 (development aid).    python tools/repro_overlap.py vol2 hf vol2 cornell"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))   # golden_fixtures
import torch
from path_tracer_python_b200.engine import Engine
from path_tracer_python_b200 import scenes, scene_compiler as SC, bvh
import golden_fixtures
e = Engine(0)
for k, v in [kv.split("=") for kv in os.environ.get("OPTS", "").split(",") if kv]:
    e.set_option(k, int(v))
def cornell():
    world, cam = scenes.cornell_config4(1024, 64); cam.initialize()
    cs = SC.compile_scene(world); cs.update(bvh.compile_bvh(cs)); cs.update(SC.perlin_tables())
    v = lambda a: [a.x, a.y, a.z]
    meta = {"width": 1024, "height": cam.img_height, "background": (0, 0, 0), "max_depth": 50,
            "camera": {"cam_center": v(cam.center), "cam_pixel00": v(cam.pixel00_loc), "cam_delta_u": v(cam.delta_u), "cam_delta_v": v(cam.delta_v),
                       "cam_defocus_disk_u": v(cam.defocus_disk_u), "cam_defocus_disk_v": v(cam.defocus_disk_v), "cam_defocus_angle": 0.0}}
    return cs, meta
order = sys.argv[1:] or ["vol2", "cornell"]
for name in order:
    if name == "vol2":
        cs, meta = golden_fixtures.load_config("vol2_final_scene", 1920, 16 / 9)
    elif name == "hf":
        cs, meta = scenes.heightfield_scene()
    else:
        cs, meta = cornell()
    e.set_scene(cs); e.set_camera(meta["camera"], meta.get("background", (0, 0, 0)), meta.get("max_depth", 50))
    W, H = meta["width"], meta["height"]
    acc = e.new_accum(W, H)
    for spp in (64, 64):
        acc.zero_(); e.render_wavefront(acc, 0, spp, 5); torch.cuda.synchronize()
        print(name, W, H, spp, "ok", float(acc.mean()), flush=True)
print("done", flush=True)
