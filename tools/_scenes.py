"""Scene selection shared by the development tools: SCENE = a `scenes.load_config` name (default
vol2_final_scene), `cornell4` (BASELINE config 4) or `heightfield` (config 3).  Returns (cs, meta, W)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))   # golden_fixtures
from path_tracer_python_b200 import scenes
import golden_fixtures


def load(W=None):
    name = os.environ.get("SCENE", "vol2_final_scene")
    if name == "cornell4":
        from path_tracer_python_b200 import scene_compiler as SC, bvh
        W = W or int(os.environ.get("W", 1024))
        world, cam = scenes.cornell_config4(W, 64); cam.initialize()
        cs = SC.compile_scene(world); cs.update(bvh.compile_bvh(cs)); cs.update(SC.perlin_tables())
        v = lambda a: [a.x, a.y, a.z]
        meta = {"height": cam.img_height, "background": (0, 0, 0), "max_depth": 50,
                "camera": {"cam_center": v(cam.center), "cam_pixel00": v(cam.pixel00_loc), "cam_delta_u": v(cam.delta_u),
                           "cam_delta_v": v(cam.delta_v), "cam_defocus_disk_u": v(cam.defocus_disk_u),
                           "cam_defocus_disk_v": v(cam.defocus_disk_v), "cam_defocus_angle": 0.0}}
        return cs, meta, W
    if name == "heightfield":
        cs, meta = scenes.heightfield_scene()
        return cs, meta, meta["width"]
    W = W or int(os.environ.get("W", 1920))
    cs, meta = golden_fixtures.load_config(name, W, 16 / 9)
    meta.setdefault("background", (0, 0, 0)); meta.setdefault("max_depth", 50)
    meta["background"] = (0, 0, 0)
    return cs, meta, W
