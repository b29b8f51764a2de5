"""TEST INFRASTRUCTURE: loaders for the golden fixtures in this directory (written by oracle/gen_golden.py
from the reference's own compile_scene / compile_bvh / kernels).  Used by tests/, bench.py's CPU arm checks,
__graft_entry__.smoke() and tools/ -- never by the product package, which builds its scenes from its own
scene classes (path_tracer_python_b200/scenes.py)."""
from __future__ import annotations

import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name: str):
    """-> (compiled_scene dict, meta dict, golden outputs dict)"""
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    g = np.load(path)
    cs, out = {}, {}
    for k in g.files:
        if k.startswith("in_"):
            cs[k[3:]] = g[k]
        elif k.startswith("out_"):
            out[k[4:]] = g[k]
    for k in [k for k in cs if k.endswith("_ref")]:
        cs[k[:-4]] = np.load(os.path.join(GOLDEN_DIR, str(cs.pop(k)) + ".npz"))["image"]
    cs["num_images"] = int(cs.get("num_images", 0))
    meta = {
        "width": int(cs.pop("img_width")), "height": int(cs.pop("img_height")), "spp": int(cs.pop("spp")),
        "background": tuple(float(x) for x in cs.pop("bg")), "max_depth": int(cs.pop("max_depth")),
        "rng_seed": int(cs.pop("rng_seed")),
    }
    cam = {k: cs.pop(k) for k in list(cs) if k.startswith("cam_")}
    meta["camera"] = cam
    return cs, meta, out



def load_config(name: str, img_width: int, aspect_ratio: float):
    """Scene arrays of a golden fixture with the camera re-targeted to another resolution."""
    cs, meta, _ = load_fixture(name)
    from path_tracer_python_b200.scenes import _CAMERAS, camera_vectors
    cam, h = camera_vectors(img_width=img_width, aspect_ratio=aspect_ratio, **_CAMERAS[name])
    meta = dict(meta, width=img_width, height=h, camera=cam)
    return cs, meta
