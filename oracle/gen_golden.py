"""TEST INFRASTRUCTURE: mint the golden fixtures under tests/golden/ from the reference itself.

Runs in THIS container only (needs /root/reference).  For each golden scene it

  1. builds the scene with the reference's own scene code (src/scenes.py or src/core classes),
  2. constructs the reference's unmodified ``TaichiRenderer`` (renderer.py:25) with the `taichi`
     module replaced by oracle/taichi_shim (pure Python, PT-F32 numerics), which runs
     compile_scene / compile_bvh / _upload_to_gpu / _upload_camera exactly as the reference does,
  3. dumps the first-hit table after ``generate_camera_rays`` + ``intersect_rays``
     (kernels.py:1219-1263) and the accumulation buffers after ``render()`` (megakernel,
     renderer.py:361) and ``render_wavefront()`` (renderer.py:249),
  4. stores inputs (compiled arrays, Perlin tables, camera) and outputs in one .npz.

The only deviation from the pristine reference is for `vol2_final_scene`: its 2401 quads exceed
``MAX_QUADS = 2048`` (fields.py:16), so that scene is run against a private temp copy of
taichi_renderer/ in which the three MAX_* capacities are raised (no kernel edits) -- the procedure
SURVEY.md section 7 prescribes.

Usage:  python oracle/gen_golden.py [scene ...]
"""
from __future__ import annotations

import contextlib
import io
import os
import random
import shutil
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLDEN = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "taichi_shim"))  # `import taichi` -> the shim

import taichi as ti  # noqa: E402  (the shim)
import refimport as R  # noqa: E402

SEED_RNG = 20261018  # device RNG seed used for all goldens


def _feature_zoo():
    """A small scene that touches every branch of kernels.py: sphere/quad/triangle hits,
    Lambert (solid, checker, noise, image), metal, dielectric, emissive, a constant medium with a
    multi-primitive boundary and one with a sphere boundary, plus defocus blur."""
    R.import_reference_host(stub_taichi=False)
    from core import Sphere, quad, triangle, hittable_list, bvh_node, camera  # type: ignore
    from core.constant_medium import constant_medium  # type: ignore
    from core.material import lambertian, metal, dielectric, diffuse_light  # type: ignore
    from core.texture import checker_texture, noise_texture, image_texture  # type: ignore
    from util import color, point3, vec3  # type: ignore
    import scenes  # type: ignore

    random.seed(77)
    world = hittable_list()
    checker = checker_texture.from_colors(0.8, color(0.2, 0.3, 0.1), color(0.9, 0.9, 0.9))
    world.add(Sphere.stationary(point3(0, -100, 0), 100, lambertian.from_texture(checker)))
    world.add(Sphere.stationary(point3(-2.2, 1.0, 0.3), 1.0, lambertian.from_texture(noise_texture(4.0))))
    cwd = os.getcwd()
    os.chdir(R.REF_SRC)
    try:
        earth = image_texture("assets/images/earthmap.jpg")
    finally:
        os.chdir(cwd)
    world.add(Sphere.stationary(point3(2.3, 1.0, -0.4), 1.0, lambertian.from_texture(earth)))
    world.add(Sphere.stationary(point3(0.0, 1.0, 0.0), 1.0, dielectric(1.5)))
    world.add(Sphere.stationary(point3(0.0, 1.0, 0.0), 0.6, metal(color(0.8, 0.6, 0.2), 0.3)))
    world.add(Sphere.moving(point3(-0.8, 0.35, 2.0), point3(-0.8, 0.6, 2.0), 0.35,
                            lambertian.from_color(color(0.7, 0.3, 0.1))))
    world.add(quad(point3(-3, 4.5, -2), vec3(6, 0, 0), vec3(0, 0, 4), diffuse_light.from_color(color(5, 5, 5))))
    world.add(quad(point3(-4, 0, -3), vec3(8, 0, 0), vec3(0, 4, 0), metal(color(0.8, 0.8, 0.9), 0.05)))
    world.add(triangle(point3(1.0, 0.0, 2.2), point3(2.2, 0.0, 1.6), point3(1.5, 1.3, 1.9),
                       lambertian.from_color(color(0.1, 0.2, 0.8))))
    world.add(triangle(point3(3.2, 0.0, 1.0), point3(3.6, 0.0, 2.2), point3(3.3, 1.6, 1.5), dielectric(1.3)))
    smoke_box = scenes.box(point3(-3.4, 0.0, 1.2), point3(-1.9, 1.2, 2.6), lambertian.from_color(color(1, 1, 1)), 20)
    world.add(constant_medium.from_color(smoke_box, color(0.9, 0.9, 0.9), 0.9))
    fog = Sphere.stationary(point3(1.2, 0.45, 2.9), 0.45, dielectric(1.5))
    world.add(fog)
    world.add(constant_medium.from_color(fog, color(0.2, 0.4, 0.9), 2.5))
    bvh = bvh_node.from_objects(world.objects, 0, len(world.objects))
    world = hittable_list()
    world.add(bvh)
    cam = camera()
    cam.aspect_ratio = 4.0 / 3.0
    cam.img_width = 32
    cam.samples_per_pixel = 6
    cam.vfov = 38
    cam.lookfrom = point3(1.5, 2.6, 9.0)
    cam.lookat = point3(0, 0.9, 0)
    cam.vup = vec3(0, 1, 0)
    cam.defocus_angle = 0.8
    cam.focus_distance = 9.3
    return world, cam


# name -> (builder, img_width, aspect, spp, bg, max_depth, raise_caps)
def _scene_specs():
    def from_scenes(name):
        return lambda: R.capture_scene(name, seed=1234)

    return {
        "wavefront_comparison": (from_scenes("wavefront_comparison"), 48, 16.0 / 9.0, 4, (0.70, 0.80, 1.00), 50, False),
        "cornell_smoke": (from_scenes("cornell_smoke"), 28, 1.0, 4, (0.0, 0.0, 0.0), 50, False),
        "feature_zoo": (_feature_zoo, 32, 4.0 / 3.0, 6, (0.55, 0.65, 0.9), 12, False),
        "vol2_final_scene": (from_scenes("vol2_final_scene"), 32, 16.0 / 9.0, 2, (0.0, 0.0, 0.0), 50, True),
    }


def _load_reference_renderer(raise_caps: bool):
    """Import render_server.taichi_renderer (pristine, or a temp copy with MAX_* raised)."""
    R.import_reference_host(stub_taichi=False)
    for m in [k for k in sys.modules if k.startswith("render_server.taichi_renderer")]:
        del sys.modules[m]
    import render_server  # type: ignore

    pkg_dir = os.path.join(R.REF_SRC, "render_server", "taichi_renderer")
    tmp = None
    if raise_caps:
        tmp = tempfile.mkdtemp(prefix="pt_ref_caps_")
        dst = os.path.join(tmp, "taichi_renderer")
        shutil.copytree(pkg_dir, dst)
        fpath = os.path.join(dst, "fields.py")
        src = open(fpath).read()
        for cap, val in (("MAX_QUADS = 2048", "MAX_QUADS = 4096"), ("MAX_BVH_NODES = 8192", "MAX_BVH_NODES = 16384")):
            assert cap in src
            src = src.replace(cap, val)
        open(fpath, "w").write(src)
        pkg_dir = dst
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "render_server.taichi_renderer", os.path.join(pkg_dir, "__init__.py"),
        submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["render_server.taichi_renderer"] = mod
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod, tmp


def generate(name: str):
    builder, width, aspect, spp, bg, max_depth, raise_caps = _scene_specs()[name]
    t_start = time.time()
    world, cam = builder()
    cam.img_width = width
    cam.aspect_ratio = aspect
    cam.samples_per_pixel = spp

    mod, tmp = _load_reference_renderer(raise_caps)
    fields = sys.modules["render_server.taichi_renderer.fields"]
    kernels = sys.modules["render_server.taichi_renderer.kernels"]
    TaichiRenderer = mod.TaichiRenderer
    out_png = os.path.join(tempfile.gettempdir(), f"golden_{name}.png")
    from util import color  # type: ignore

    def new_renderer():
        with contextlib.redirect_stdout(io.StringIO()):
            r = TaichiRenderer(world, cam, out_png)
        r.background_color = color(*bg)
        r.max_depth = max_depth
        fields.ray_pixel_index.on_read = ti.runtime.set_pixel
        ti.runtime.image_width = cam.img_width
        ti.runtime.seed = SEED_RNG
        ti.runtime.sample_counter = 0
        return r

    gold = {}
    # ---- inputs: what the reference compiled and uploaded
    r = new_renderer()
    W, H = int(cam.img_width), int(cam.img_height)
    cs = R.reference_compile(world)
    cs["perlin_randvec"] = fields.perlin_randvec.to_numpy()
    cs["perlin_perm_x"] = fields.perlin_perm_x.to_numpy()
    cs["perlin_perm_y"] = fields.perlin_perm_y.to_numpy()
    cs["perlin_perm_z"] = fields.perlin_perm_z.to_numpy()
    # make sure what sits in the (shim) device fields is what reference_compile returned
    ns, nn = cs["sphere_data"].shape[0], cs["bvh_bbox_min"].shape[0]
    assert np.array_equal(fields.sphere_data.to_numpy()[:ns], cs["sphere_data"])
    assert np.array_equal(fields.bvh_nodes.members["bbox_min"].to_numpy()[:nn], cs["bvh_bbox_min"])
    assert int(fields.num_bvh_nodes[None]) == nn
    images = {k: cs.pop(k) for k in [k for k in cs if k.startswith("image_")]}
    for k, v in cs.items():
        gold["in_" + k] = v
    for k, v in images.items():  # textures live in their own fixture, referenced by shape
        assert v.shape == (512, 1024, 3)
        gold["in_" + k + "_ref"] = np.array("earthmap_u8")
    for k, v in R.camera_arrays(cam).items():
        gold["in_" + k] = v
    gold["in_bg"] = np.array(bg, np.float64)
    gold["in_max_depth"] = np.int32(max_depth)
    gold["in_spp"] = np.int32(spp)
    gold["in_rng_seed"] = np.uint32(SEED_RNG)

    # ---- first hit of sample 0's camera rays (wavefront stage outputs)
    r._upload_camera()
    kernels.generate_camera_rays(W, H)
    kernels.intersect_rays()
    n = W * H
    gold["out_ray_o"] = fields.ray_origins.to_numpy()[:n]
    gold["out_ray_d"] = fields.ray_directions.to_numpy()[:n]
    gold["out_hit_valid"] = fields.hit_valid.to_numpy()[:n]
    gold["out_hit_t"] = fields.hit_t.to_numpy()[:n]
    gold["out_hit_prim_type"] = fields.hit_prim_type.to_numpy()[:n]
    gold["out_hit_prim_idx"] = fields.hit_prim_idx.to_numpy()[:n]
    gold["out_hit_point"] = fields.hit_point.to_numpy()[:n]
    gold["out_hit_normal"] = fields.hit_normal.to_numpy()[:n]
    kernels.clear_accum_buffer()
    t_fh = time.time()

    # ---- megakernel render(): the reference's own loop
    with contextlib.redirect_stdout(io.StringIO()):
        r.render(enable_preview=False)
    gold["out_accum_megakernel"] = fields.accum_buffer.to_numpy()
    draws_mega = ti.runtime.draws
    t_mega = time.time()

    # ---- wavefront render_wavefront()
    r = new_renderer()
    # every TaichiRenderer.__init__ draws fresh Perlin tables from the global `random` stream
    # (renderer.py:78-80); re-upload the first renderer's tables so both modes see one scene
    fields.perlin_randvec.from_numpy(cs["perlin_randvec"])
    fields.perlin_perm_x.from_numpy(cs["perlin_perm_x"])
    fields.perlin_perm_y.from_numpy(cs["perlin_perm_y"])
    fields.perlin_perm_z.from_numpy(cs["perlin_perm_z"])
    with contextlib.redirect_stdout(io.StringIO()):
        r.render_wavefront(enable_preview=False)
    gold["out_accum_wavefront"] = fields.accum_buffer.to_numpy()
    t_wave = time.time()

    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, f"{name}.npz")
    np.savez_compressed(path, **gold)
    if images:
        np.savez_compressed(os.path.join(GOLDEN, "earthmap_u8.npz"), image=list(images.values())[0])
    if tmp:
        shutil.rmtree(tmp, ignore_errors=True)
    print(f"{name}: {W}x{H} spp={spp} prims={ns}s/{cs['quad_Q'].shape[0]}q/{cs['triangle_v0'].shape[0]}t "
          f"nodes={nn} first-hit {t_fh - t_start:.1f}s mega {t_mega - t_fh:.1f}s wave {t_wave - t_mega:.1f}s "
          f"draws={draws_mega} -> {os.path.relpath(path, REPO)} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    names = sys.argv[1:] or list(_scene_specs())
    for nm in names:
        generate(nm)
