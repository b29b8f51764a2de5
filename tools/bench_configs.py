"""Secondary benchmark lines for BASELINE.json configs 1, 3, 4, 5 (config 2 is bench.py's headline).
Prints one JSON line per measurement; run on one GPU."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))   # golden_fixtures
import torch
from path_tracer_python_b200.engine import Engine, algorithmic_bytes
from path_tracer_python_b200 import scenes, scene_compiler as SC, bvh
import golden_fixtures

e = Engine(0)

def timed(fn, *a):
    fn(*a); torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.record(); fn(*a); y.record(); torch.cuda.synchronize()
    return x.elapsed_time(y)

def render_line(tag, cs, meta, spp, seed=1234, flight=None):
    e.set_scene(cs); e.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    W, H = meta["width"], meta["height"]
    out = {"config": tag, "width": W, "height": H, "spp": spp}
    for mode, fn in (("megakernel", e.render_megakernel), ("wavefront", e.render_wavefront)):
        acc = e.new_accum(W, H)
        ms = timed(lambda: (acc.zero_(), fn(acc, 0, spp, seed)))
        out[mode + "_mpaths_s"] = round(W * H * spp / ms / 1e3, 1)
        out[mode + "_ms"] = round(ms, 2)
    e.enable_counters(True); e.counters(reset=True)
    acc = e.new_accum(W, H); e.render_megakernel(acc, 0, min(spp, 4), seed); c = e.counters(); e.enable_counters(False)
    out["rays_per_path"] = round(c["rays"] / c["paths"], 3); out["box_tests_per_ray"] = round(c["box_tests"] / c["rays"], 2)
    out["alg_bytes_per_path"] = round(algorithmic_bytes(c) / c["paths"], 1)
    print(json.dumps(out), flush=True)

# config 1: wavefront_comparison 400x225 x 16 spp
cs, meta = golden_fixtures.load_config("wavefront_comparison", 400, 16 / 9); meta["background"] = (0.7, 0.8, 1.0); meta["max_depth"] = 50
render_line("1: wavefront_comparison 400x225 16spp", cs, meta, 16)
render_line("1b: wavefront_comparison 1920x1080 64spp", *golden_fixtures.load_config("wavefront_comparison", 1920, 16 / 9), 64)

# config 3: 1 M triangles, primary rays + 8-bounce paths
t0 = time.time(); cs, meta = scenes.heightfield_scene(); build = time.time() - t0
info = e.set_scene(cs); e.set_camera(meta["camera"], meta["background"], meta["max_depth"])
W, H = meta["width"], meta["height"]
o, d = e.generate_camera_rays(W, H, 7, 0)
ms = timed(lambda: e.first_hit(o, d, want_pn=False))
e.enable_counters(True); e.counters(reset=True); e.first_hit(o, d, want_pn=False); c = e.counters(); e.enable_counters(False)
print(json.dumps({"config": "3: 1M-triangle height field, primary rays 1920x1080", "scene_build_s": round(build, 2), "nodes": len(cs["bvh_prim_idx"]),
                  "tree_depth": info.max_depth_tree, "scene_bytes": int(info.total_bytes), "primary_mrays_s": round(W * H / ms / 1e3, 1), "ms": round(ms, 3),
                  "box_tests_per_ray": round(c["box_tests"] / c["rays"], 2), "tri_tests_per_ray": round(c["triangle_tests"] / c["rays"], 2),
                  "alg_gb_s": round(algorithmic_bytes(c) / ms / 1e6, 1)}), flush=True)
render_line("3: 1M-triangle height field, 8-bounce paths 1920x1080 16spp", cs, meta, 16)

# config 4: Cornell + medium + glass + SSS-as-Lambert, 1024^2
world, cam = scenes.cornell_config4(1024, 64); cam.initialize()
cs = SC.compile_scene(world); cs.update(bvh.compile_bvh(cs)); cs.update(SC.perlin_tables())
v = lambda a: [a.x, a.y, a.z]
meta = {"width": 1024, "height": cam.img_height, "background": (0, 0, 0), "max_depth": 50,
        "camera": {"cam_center": v(cam.center), "cam_pixel00": v(cam.pixel00_loc), "cam_delta_u": v(cam.delta_u), "cam_delta_v": v(cam.delta_v),
                   "cam_defocus_disk_u": v(cam.defocus_disk_u), "cam_defocus_disk_v": v(cam.defocus_disk_v), "cam_defocus_angle": 0.0}}
render_line("4: Cornell+medium+glass+SSS-as-Lambert 1024x1024 4096spp (the full BASELINE config)", cs, meta, 4096)

# config 5 (one GPU's share at N = 8): vol2_final_scene 3840x2160, 128 spp -- bench.py times the whole 1024 spp at every N
render_line("5: vol2_final_scene 3840x2160 128spp (the 1/8 share of 1024)", *golden_fixtures.load_config("vol2_final_scene", 3840, 16 / 9), 128)
