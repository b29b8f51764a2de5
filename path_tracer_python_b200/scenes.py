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


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs 3 and 4: built here (the reference ships no OBJ asset and caps triangles
# at 4096), with the reference's own float64 -> float32 conventions.
# ---------------------------------------------------------------------------------------------
def _material_block(prefix, n, mat_type=0, color=(0.8, 0.8, 0.8), emit=(0, 0, 0), fuzz=0.0, ir=1.0):
    c = np.tile(np.asarray(color, np.float32), (n, 1))
    return {
        f"{prefix}_material_type": np.full(n, mat_type, np.int32), f"{prefix}_material_albedo": c.copy(),
        f"{prefix}_material_fuzz": np.full(n, fuzz, np.float32), f"{prefix}_material_ir": np.full(n, ir, np.float32),
        f"{prefix}_material_emit_color": np.tile(np.asarray(emit, np.float32), (n, 1)),
        f"{prefix}_texture_type": np.zeros(n, np.int32), f"{prefix}_texture_scale": np.ones(n, np.float32),
        f"{prefix}_texture_color1": c.copy(), f"{prefix}_texture_color2": c.copy(),
        f"{prefix}_texture_image_idx": np.full(n, -1, np.int32), f"{prefix}_is_constant_medium": np.zeros(n, np.int32),
        f"{prefix}_medium_density": np.zeros(n, np.float32), f"{prefix}_medium_albedo": np.zeros((n, 3), np.float32),
    }


def heightfield_scene(nx: int = 1000, ny: int = 500, seed: int = 7, img_width: int = 1920, aspect_ratio: float = 16 / 9):
    """Config 3 stand-in: a seeded height field of nx*ny cells = 2*nx*ny triangles (1 M by default),
    Lambert (0.6, 0.4, 0.2), over a ground sphere, lit by one emissive quad; max_depth 8.
    Triangle records follow core/triangle.py:14-33 (edges and unit normal in float64, then narrowed
    to float32 like compile_triangle_geometry, scene_compiler.py:214-251)."""
    from . import bvh as _bvh

    rng = np.random.default_rng(seed)
    xs = np.linspace(-10.0, 10.0, nx + 1)
    ys = np.linspace(-5.0, 5.0, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    Z = np.zeros_like(X)
    for k in range(1, 5):
        Z += (0.15 / k) * np.sin(0.7 * k * X + 0.3 * k) * np.sin(0.9 * k * Y + 1.1 * k)
    Z += rng.uniform(-0.01, 0.01, Z.shape)
    P = np.stack([X, Z, Y], axis=-1)                      # y is up
    p00, p10, p01, p11 = P[:-1, :-1], P[:-1, 1:], P[1:, :-1], P[1:, 1:]
    v0 = np.concatenate([p00.reshape(-1, 3), p11.reshape(-1, 3)])
    v1 = np.concatenate([p01.reshape(-1, 3), p10.reshape(-1, 3)])
    v2 = np.concatenate([p10.reshape(-1, 3), p01.reshape(-1, 3)])
    e1, e2 = v1 - v0, v2 - v0
    nrm = np.cross(e1, e2)
    nrm /= np.sqrt((nrm ** 2).sum(axis=1, keepdims=True))
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    cs = {"triangle_v0": f32(v0), "triangle_v1": f32(v1), "triangle_v2": f32(v2), "triangle_edge1": f32(e1),
          "triangle_edge2": f32(e2), "triangle_normal": f32(nrm)}
    cs.update(_material_block("triangle", len(v0), 0, (0.6, 0.4, 0.2)))
    cs["sphere_data"] = np.array([[0.0, -1000.6, 0.0, 1000.0]], np.float32)
    cs.update(_material_block("sphere", 1, 0, (0.5, 0.5, 0.5)))
    Q, u, v = np.array([-4.0, 6.0, -3.0]), np.array([8.0, 0.0, 0.0]), np.array([0.0, 0.0, 6.0])
    n = np.cross(u, v)
    unit = n / np.sqrt(n @ n)
    cs.update({"quad_Q": f32([Q]), "quad_u": f32([u]), "quad_v": f32([v]), "quad_normal": f32([unit]),
               "quad_D": f32([unit @ Q]), "quad_w": f32([n / (n @ n)])})
    cs.update(_material_block("quad", 1, 3, (0, 0, 0), emit=(6, 6, 6)))
    cs["num_images"] = 0
    cs.update(_bvh.compile_bvh(cs))
    cam, h = camera_vectors((0.0, 7.0, 13.0), (0.0, 0.0, 0.0), (0, 1, 0), 40, img_width, aspect_ratio)
    meta = {"width": img_width, "height": h, "camera": cam, "background": (0.6, 0.7, 0.9), "max_depth": 8}
    return cs, meta


def cornell_config4(img_width: int = 1024, spp: int = 4096):
    """Config 4: Cornell box (src/scenes.py:1031-1041, light 15) + the tall box as a constant medium
    (density 0.01, white; :1114-1124) + a glass sphere r=90 at (190,90,190) + a subsurface sphere
    r=80 at (400,80,150) -- which the Taichi path renders as 0.8-grey Lambert
    (scene_compiler.py:406-417).  Returns (world, camera) built from path_tracer_python_b200.scene."""
    from . import scene as sc

    w = sc.hittable_list()
    red, white, green = (sc.lambertian.from_color(sc.color(*c)) for c in ((.65, .05, .05), (.73, .73, .73), (.12, .45, .15)))
    light = sc.diffuse_light.from_color(sc.color(15, 15, 15))
    P, V = sc.point3, sc.vec3
    w.add(sc.quad(P(555, 0, 0), V(0, 0, 555), V(0, 555, 0), green))
    w.add(sc.quad(P(0, 0, 0), V(0, 555, 0), V(0, 0, 555), red))
    w.add(sc.quad(P(343, 554, 332), V(-130, 0, 0), V(0, 0, -105), light))
    w.add(sc.quad(P(0, 0, 0), V(0, 0, 555), V(555, 0, 0), white))
    w.add(sc.quad(P(555, 555, 555), V(-555, 0, 0), V(0, 0, -555), white))
    w.add(sc.quad(P(0, 0, 555), V(0, 555, 0), V(555, 0, 0), white))
    w.add(sc.constant_medium.from_color(sc.box(P(265, 0, 295), P(430, 330, 460), white, 15), sc.color(1, 1, 1), 0.01))
    w.add(sc.Sphere.stationary(P(190, 90, 190), 90, sc.dielectric(1.5)))
    w.add(sc.Sphere.stationary(P(400, 80, 150), 80, sc.subsurface_volumetric(sc.color(.2, .5, .2), 0.08, 0.8, 0.7)))
    cam = sc.camera()
    cam.aspect_ratio, cam.img_width, cam.samples_per_pixel = 1.0, img_width, spp
    cam.vfov, cam.lookfrom, cam.lookat, cam.vup = 40, P(278, 278, -800), P(278, 278, 0), V(0, 1, 0)
    cam.defocus_angle = 0
    return w, cam
