"""Benchmark / test scenes as compiled arrays.

`/root/reference` (and therefore src/scenes.py + src/core) does not exist on the GPU box, so the
scenes BASELINE.json names are shipped as the arrays the reference's own `compile_scene` +
`compile_bvh` produced for them (tests/golden/*.npz, written by oracle/gen_golden.py, seed 1234).
Resolution / spp / depth are re-targeted here by recomputing the camera vectors with the
reference's formulas (src/core/camera.py:34-72).
"""
from __future__ import annotations

import math
import os

import numpy as np

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(REPO_ROOT, "tests", "golden")


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


def camera_vectors(lookfrom, lookat, vup, vfov, img_width, aspect_ratio, defocus_angle=0.0, focus_distance=10.0):
    """The seven vectors core/camera.py:34-72 `initialize()` derives, in float64, with the
    reference's operation order (util/vec3.py: length = sqrt(x**2+y**2+z**2), normalise = v / len)."""
    def sub(a, b): return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]
    def add(a, b): return [a[0] + b[0], a[1] + b[1], a[2] + b[2]]
    def mul(a, s): return [a[0] * s, a[1] * s, a[2] * s]
    def div(a, s): return [a[0] / s, a[1] / s, a[2] / s]
    def norm(a): return div(a, math.sqrt(a[0] ** 2 + a[1] ** 2 + a[2] ** 2))
    def cross(a, b): return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]

    lookfrom, lookat, vup = ([float(x) for x in v] for v in (lookfrom, lookat, vup))
    img_height = 1 if int(img_width / aspect_ratio) < 1 else int(img_width / aspect_ratio)
    theta = vfov * math.pi / 180.0
    h = math.tan(theta / 2)
    viewport_height = 2.0 * h * focus_distance
    viewport_width = viewport_height * (img_width / img_height)
    w = norm(sub(lookfrom, lookat))
    u = norm(cross(vup, w))
    v = cross(w, u)
    viewport_u = mul(u, viewport_width)
    viewport_v = mul([-v[0], -v[1], -v[2]], viewport_height)
    delta_u = div(viewport_u, img_width)
    delta_v = div(viewport_v, img_height)
    upper_left = sub(sub(sub(lookfrom, mul(w, focus_distance)), div(viewport_u, 2)), div(viewport_v, 2))
    pixel00 = add(upper_left, mul(add(delta_u, delta_v), 0.5))
    defocus_radius = focus_distance * math.tan((defocus_angle * math.pi / 180.0) / 2)
    f = lambda a: np.array(a, dtype=np.float64)
    return {
        "cam_center": f(lookfrom), "cam_pixel00": f(pixel00), "cam_delta_u": f(delta_u), "cam_delta_v": f(delta_v),
        "cam_defocus_disk_u": f(mul(u, defocus_radius)), "cam_defocus_disk_v": f(mul(v, defocus_radius)),
        "cam_defocus_angle": np.float64(defocus_angle),
    }, img_height


# camera definitions of the BASELINE.json configs (src/scenes.py:1231-1237 and :1497-1503)
_CAMERAS = {
    "vol2_final_scene": dict(lookfrom=(478, 278, -600), lookat=(278, 278, 0), vup=(0, 1, 0), vfov=40),
    "wavefront_comparison": dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vup=(0, 1, 0), vfov=20),
    "cornell_smoke": dict(lookfrom=(278, 278, -800), lookat=(278, 278, 0), vup=(0, 1, 0), vfov=40),
}


def load_config(name: str, img_width: int, aspect_ratio: float):
    """Scene arrays of a golden fixture with the camera re-targeted to another resolution."""
    cs, meta, _ = load_fixture(name)
    cam, h = camera_vectors(img_width=img_width, aspect_ratio=aspect_ratio, **_CAMERAS[name])
    meta = dict(meta, width=img_width, height=h, camera=cam)
    return cs, meta
