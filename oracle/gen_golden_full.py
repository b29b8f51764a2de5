"""TEST INFRASTRUCTURE: fixture for the *full* feature mode (moving spheres, subsurface materials).

Those features exist only in the reference's pure-Python tracer (src/core/camera.py:74-126 ray_color,
core/sphere.py:35, core/material.py:144-276) -- the Taichi path drops them -- so there are no bits to
compare; the oracle's full mode is pinned *statistically*: this script renders a small scene with
the reference's own `camera.ray_color` loop (linear radiance, no tonemap) and stores the image next
to the scene arrays.  tests/test_full_features.py compares region means.

Runs in THIS container only (needs /root/reference).  ~2 minutes.
"""
import os, random, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE); sys.path.insert(0, REPO)
import refimport as R

def build():
    R.import_reference_host()
    from core import Sphere, quad, hittable_list, camera  # type: ignore
    from core.material import lambertian, diffuse_light, subsurface_simple, subsurface_volumetric  # type: ignore
    from util import color, point3, vec3  # type: ignore
    world = hittable_list()
    world.add(Sphere.stationary(point3(0, -1000, 0), 1000, lambertian.from_color(color(0.5, 0.5, 0.5))))
    world.add(quad(point3(-2, 3.0, -1), vec3(4, 0, 0), vec3(0, 0, 2), diffuse_light.from_color(color(6, 6, 6))))
    world.add(Sphere.moving(point3(-1.6, 0.5, 0.0), point3(-1.6, 1.1, 0.0), 0.5, lambertian.from_color(color(0.7, 0.3, 0.1))))
    world.add(Sphere.stationary(point3(0.0, 0.5, 0.0), 0.5, subsurface_volumetric(color(0.2, 0.5, 0.2), 0.08, 0.8, 0.7)))
    world.add(Sphere.stationary(point3(1.6, 0.5, 0.0), 0.5, subsurface_simple(color(0.8, 0.6, 0.3), 0.15)))
    cam = camera()
    cam.aspect_ratio, cam.img_width, cam.samples_per_pixel = 2.0, 40, 384
    cam.vfov, cam.lookfrom, cam.lookat, cam.vup = 30, point3(0, 1.6, -6.5), point3(0, 0.6, 0), vec3(0, 1, 0)
    cam.defocus_angle = 0.0
    cam.max_depth = 12
    cam.background = color(0.70, 0.80, 1.00)
    cam.russian_roulette_enabled = False
    cam.russian_roulette_min_depth = 3
    cam.russian_roulette_threshold = 0.1
    return world, cam

if __name__ == "__main__":
    random.seed(4242)
    world, cam = build()
    cam.initialize()
    W, H, spp = cam.img_width, cam.img_height, cam.samples_per_pixel
    img = np.zeros((H, W, 3), np.float64)
    t0 = time.time()
    for h in range(H):
        for w in range(W):
            acc = [0.0, 0.0, 0.0]
            for _ in range(spp):
                c = cam.ray_color(cam.get_ray(w, h), cam.max_depth, world)
                acc[0] += c.x; acc[1] += c.y; acc[2] += c.z
            img[h, w] = [a / spp for a in acc]
        print(f"row {h + 1}/{H} {time.time() - t0:.0f}s", end="\r", flush=True)
    from path_tracer_python_b200 import scene_compiler as SC, bvh
    cs = SC.compile_scene(world, features="full")
    cs.update(bvh.compile_bvh(cs))
    out = {"in_" + k: np.asarray(v) for k, v in cs.items()}
    for k, v in R.camera_arrays(cam).items():
        out["in_" + k] = v
    out.update({"in_bg": np.array([0.7, 0.8, 1.0]), "in_max_depth": np.int32(12), "in_spp": np.int32(spp), "in_rng_seed": np.uint32(7),
                "out_radiance_pure_python": img.astype(np.float32)})
    path = os.path.join(REPO, "tests", "golden", "full_features.npz")
    np.savez_compressed(path, **out)
    print(f"\nwrote {path} ({os.path.getsize(path) // 1024} KiB) in {time.time() - t0:.0f}s; mean radiance {img.mean():.4f}")
