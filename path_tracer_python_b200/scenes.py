"""The scenes BASELINE.json names, built with this package's own scene classes (scene.py) and compiled by
its own compiler + native SAH builder -- `/root/reference` (src/scenes.py, src/core) does not exist on the GPU
box.  Scene definitions follow the reference's (src/scenes.py, cited per function): same objects, same
parameters, same draws from Python's `random` stream, so seed 1234 yields the same geometry.  Camera vectors
come from the reference's formulas (src/core/camera.py:34-72).  The only asset is the earth texture
(data/earthmap_u8.npz).  (Golden fixtures -- the reference's own compiled arrays and outputs -- live under
tests/golden and are read by tests/golden_fixtures.py, never by this package.)
"""
from __future__ import annotations

import math
import os

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def earth_texture() -> np.ndarray:
    """The earth map the reference's scenes use (src/assets/images/earthmap.jpg) as 8-bit RGB, shipped as
    package data: texel = u8 / 255.0f reproduces the reference's float32 image bit for bit."""
    return np.load(os.path.join(DATA_DIR, "earthmap_u8.npz"))["image"]


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


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs 1 and 2 / 5, through the scene classes.
# ---------------------------------------------------------------------------------------------
def vol2_final_scene(img_width: int = 1920, aspect_ratio: float = 16 / 9, spp: int = 256, seed: int | None = 1234):
    """Configs 2 and 5: `vol2_final_scene` (src/scenes.py:1152-1244): 20 x 20 ground boxes of random height
    (2400 quads), a light quad, a moving / a glass / a metal sphere, a glass sphere doubling as a medium
    boundary, the r = 5000 'fog' sphere around everything, the earth, a Perlin sphere and 1000 small spheres
    in a cube.  Returns (world, camera); `seed` re-seeds Python's `random` first (the reference's scenes draw
    from the global stream, and so do the Perlin tables the renderer makes afterwards)."""
    import random

    from . import scene as sc
    if seed is not None:
        random.seed(seed)
    P, V, C = sc.point3, sc.vec3, sc.color
    ground = sc.lambertian.from_color(C(0.48, 0.83, 0.53))
    boxes = sc.hittable_list()
    for i in range(20):
        for j in range(20):
            x0, z0 = -1000.0 + i * 100.0, -1000.0 + j * 100.0
            boxes.add(sc.box(P(x0, 0.0, z0), P(x0 + 100.0, random.uniform(1, 101), z0 + 100.0), ground))
    world = sc.hittable_list()
    world.add(sc.bvh_node.from_objects(boxes.objects, 0, len(boxes.objects)))
    world.add(sc.quad(P(123, 554, 147), V(300, 0, 0), V(0, 0, 265), sc.diffuse_light.from_color(C(7, 7, 7))))
    c1 = P(400, 400, 200)
    world.add(sc.Sphere.moving(c1, c1 + V(30, 0, 0), 50, sc.lambertian.from_color(C(0.7, 0.3, 0.1))))
    world.add(sc.Sphere.stationary(P(260, 150, 45), 50, sc.dielectric(1.5)))
    world.add(sc.Sphere.stationary(P(0, 150, 145), 50, sc.metal(C(0.8, 0.8, 0.9), 1.0)))
    boundary = sc.Sphere.stationary(P(360, 150, 145), 70, sc.dielectric(1.5))
    world.add(boundary)
    world.add(sc.constant_medium.from_color(boundary, C(0.2, 0.4, 0.9), 0.2))
    world.add(sc.constant_medium.from_color(sc.Sphere.stationary(P(0, 0, 0), 5000, sc.dielectric(1.5)), C(1, 1, 1), 0.0001))
    world.add(sc.Sphere.stationary(P(400, 200, 400), 100, sc.lambertian.from_texture(sc.image_texture(earth_texture()))))
    world.add(sc.Sphere.stationary(P(220, 280, 300), 80, sc.lambertian.from_texture(sc.noise_texture(0.2))))
    white = sc.lambertian.from_color(C(0.73, 0.73, 0.73))
    cube = sc.hittable_list()
    for _ in range(1000):
        cube.add(sc.Sphere.stationary(P(random.uniform(0, 165), random.uniform(0, 165), random.uniform(0, 165)) + V(-100, 270, 395), 10, white))
    world.add(sc.bvh_node.from_objects(cube.objects, 0, len(cube.objects)))
    top = sc.hittable_list([sc.bvh_node.from_objects(world.objects, 0, len(world.objects))])
    cam = sc.camera()
    cam.aspect_ratio, cam.img_width, cam.samples_per_pixel = aspect_ratio, img_width, spp
    c = _CAMERAS["vol2_final_scene"]
    cam.vfov, cam.lookfrom, cam.lookat, cam.vup, cam.defocus_angle = c["vfov"], P(*c["lookfrom"]), P(*c["lookat"]), V(*c["vup"]), 0
    return top, cam


def wavefront_comparison(img_width: int = 400, aspect_ratio: float = 16 / 9, spp: int = 16, seed: int | None = 1234):
    """Config 1: `wavefront_comparison` (src/scenes.py:1433-1503): a ground sphere, a 6 x 6 grid of small
    diffuse / metal / glass spheres, three large ones and a spherical light.  Returns (world, camera)."""
    import random

    from . import scene as sc
    if seed is not None:
        random.seed(seed)
    P, V, C = sc.point3, sc.vec3, sc.color
    rnd3 = lambda lo=0.0, hi=1.0: C(random.uniform(lo, hi), random.uniform(lo, hi), random.uniform(lo, hi))
    world = sc.hittable_list()
    world.add(sc.Sphere.stationary(P(0, -1000, 0), 1000, sc.lambertian.from_color(C(0.5, 0.5, 0.5))))
    for a in range(-3, 3):
        for b in range(-3, 3):
            choose = random.random()
            center = P(a + 0.9 * random.random(), 0.2, b + 0.9 * random.random())
            if (center - P(4, 0.2, 0)).length() > 0.9:
                if choose < 0.6:
                    c1, c2 = rnd3(), rnd3()
                    world.add(sc.Sphere.stationary(center, 0.2, sc.lambertian.from_color(C(c1.x * c2.x, c1.y * c2.y, c1.z * c2.z))))
                elif choose < 0.85:
                    albedo = rnd3(0.5, 1)
                    world.add(sc.Sphere.stationary(center, 0.2, sc.metal(albedo, random.uniform(0, 0.5))))
                else:
                    world.add(sc.Sphere.stationary(center, 0.2, sc.dielectric(1.5)))
    world.add(sc.Sphere.stationary(P(0, 1, 0), 1.0, sc.dielectric(1.5)))
    world.add(sc.Sphere.stationary(P(-4, 1, 0), 1.0, sc.lambertian.from_color(C(0.4, 0.2, 0.1))))
    world.add(sc.Sphere.stationary(P(4, 1, 0), 1.0, sc.metal(C(0.7, 0.6, 0.5), 0.0)))
    world.add(sc.Sphere.stationary(P(0, 5, 0), 1.5, sc.diffuse_light.from_color(C(4, 4, 4))))
    top = sc.hittable_list([sc.bvh_node.from_objects(world.objects, 0, len(world.objects))])
    cam = sc.camera()
    cam.aspect_ratio, cam.img_width, cam.samples_per_pixel = aspect_ratio, img_width, spp
    c = _CAMERAS["wavefront_comparison"]
    cam.vfov, cam.lookfrom, cam.lookat, cam.vup, cam.defocus_angle = c["vfov"], P(*c["lookfrom"]), P(*c["lookat"]), V(*c["vup"]), 0.0
    return top, cam


def camera_dict(cam) -> dict:
    """The seven camera vectors of an initialised scene.camera as the dict Engine.set_camera takes."""
    cam.initialize()
    v = lambda a: np.array([a.x, a.y, a.z], np.float64)
    return {"cam_center": v(cam.center), "cam_pixel00": v(cam.pixel00_loc), "cam_delta_u": v(cam.delta_u),
            "cam_delta_v": v(cam.delta_v), "cam_defocus_disk_u": v(cam.defocus_disk_u),
            "cam_defocus_disk_v": v(cam.defocus_disk_v), "cam_defocus_angle": np.float64(cam.defocus_angle)}


def compile_world(world, features: str = "taichi_compat", with_perlin: bool = True) -> dict:
    """scene graph -> the flat compiled-scene dict the engine consumes: Perlin tables, compile_scene, native SAH
    build -- in the order TaichiRenderer.__init__ makes them (the Perlin tables draw from `random` first)."""
    from . import bvh as _bvh
    from . import scene_compiler as _sc
    perlin = _sc.perlin_tables() if with_perlin else {}
    cs = _sc.compile_scene(world, features)
    cs.update(_bvh.compile_bvh(cs))
    for k, v in perlin.items():
        cs.setdefault(k, v)
    return cs
