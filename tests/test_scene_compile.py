"""Host compile: my scene classes + compiler against hand-computed values, and (when the reference is
mounted) array-for-array against the reference's own compile_scene / compile_bvh / perlin()."""
import os
import random
import sys

import numpy as np
import pytest

from path_tracer_python_b200 import bvh, scene as sc, scene_compiler as SC

REF = os.path.isdir("/root/reference/src/render_server/taichi_renderer")


def _mini_world():
    w = sc.hittable_list()
    w.add(sc.Sphere.stationary(sc.point3(0, -100, 0), 100, sc.lambertian.from_texture(
        sc.checker_texture.from_colors(0.5, sc.color(0.1, 0.2, 0.3), sc.color(0.9, 0.8, 0.7)))))
    w.add(sc.Sphere.moving(sc.point3(1, 1, 1), sc.point3(2, 1, 1), 0.5, sc.metal(sc.color(0.8, 0.6, 0.2), 1.7)))
    glass = sc.Sphere.stationary(sc.point3(0, 1, 0), 1.0, sc.dielectric(1.5))
    w.add(glass)
    w.add(sc.constant_medium.from_color(glass, sc.color(0.2, 0.4, 0.9), 0.25))
    w.add(sc.quad(sc.point3(-1, 3, -1), sc.vec3(2, 0, 0), sc.vec3(0, 0, 2), sc.diffuse_light.from_color(sc.color(4, 5, 6))))
    w.add(sc.triangle(sc.point3(0, 0, 2), sc.point3(1, 0, 2), sc.point3(0, 1, 2), sc.lambertian.from_texture(sc.noise_texture(3.0))))
    w.add(sc.constant_medium.from_color(sc.box(sc.point3(3, 0, 0), sc.point3(4, 1, 1), sc.lambertian.from_color(sc.color(1, 1, 1))),
                                        sc.color(1, 1, 1), 0.01))
    w.add(sc.Sphere.stationary(sc.point3(5, 5, 5), 1.0, sc.subsurface_volumetric(sc.color(0.2, 0.5, 0.2))))
    return w


def test_compile_own_classes():
    cs = SC.compile_scene(_mini_world())
    assert cs["sphere_data"].shape == (4, 4) and cs["quad_Q"].shape == (7, 3) and cs["triangle_v0"].shape == (1, 3)
    assert cs["sphere_data"].dtype == np.float32
    assert cs["sphere_material_type"].tolist() == [0, 1, 2, 0]           # subsurface -> grey Lambert
    assert np.allclose(cs["sphere_material_albedo"][3], 0.8) and np.allclose(cs["sphere_texture_color1"][3], 0.8)
    assert cs["sphere_data"][1].tolist() == [1.0, 1.0, 1.0, 0.5]         # centre baked at time 0
    assert cs["sphere_material_fuzz"][1] == 1.0                          # fuzz clamped to 1
    assert cs["sphere_texture_type"][0] == 1 and cs["sphere_texture_scale"][0] == np.float32(0.5)
    assert cs["sphere_is_constant_medium"].tolist() == [0, 0, 1, 0]
    assert cs["sphere_medium_density"][2] == np.float32(0.25)
    assert np.allclose(cs["sphere_medium_albedo"][2], [0.2, 0.4, 0.9])
    assert cs["quad_material_type"][0] == 3 and cs["quad_material_emit_color"][0].tolist() == [4.0, 5.0, 6.0]
    assert cs["quad_is_constant_medium"].tolist() == [0, 1, 1, 1, 1, 1, 1]
    assert np.allclose(cs["quad_medium_density"][1:], 0.01)
    assert cs["triangle_texture_type"][0] == 3 and cs["triangle_texture_scale"][0] == 3.0
    assert np.allclose(cs["quad_normal"][0], [0, -1, 0]) and np.isclose(cs["quad_D"][0], -3.0)
    tree = bvh.compile_bvh(cs)
    assert len(tree["bvh_prim_idx"]) == 2 * 12 - 1


def test_empty_world():
    cs = SC.compile_scene(sc.hittable_list())
    assert cs["sphere_data"].shape == (0, 4) and cs["quad_D"].shape == (0,) and cs["num_images"] == 0
    assert len(bvh.compile_bvh(cs)["bvh_prim_idx"]) == 0


def test_image_texture_roundtrip():
    img = (np.arange(4 * 6 * 3) % 256).astype(np.uint8).reshape(4, 6, 3)
    w = sc.hittable_list([sc.Sphere.stationary(sc.point3(0, 0, 0), 1, sc.lambertian.from_texture(sc.image_texture(img)))])
    cs = SC.compile_scene(w)
    assert cs["num_images"] == 1 and np.array_equal(cs["image_0"], img)
    assert cs["sphere_texture_type"][0] == 2 and cs["sphere_texture_image_idx"][0] == 0


def test_obj_mesh_loader_and_array_fast_path(tmp_path):
    """core/mesh.py semantics without pywavefront: scale then offset, fan triangulation, negative
    and v/vt/vn indices, degenerate faces dropped; array-backed compile == per-object compile."""
    (tmp_path / "m").mkdir()
    (tmp_path / "m" / "cube.obj").write_text(
        "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\nv 1 0 1\nv 1 1 1\nv 0 1 1\n"
        "f 1 2 3 4\nf 5/1/1 6/2/1 7/3/1 8/4/1\nf -8 -7 -3\nf 1 1 2\n")
    m = sc.mesh(str(tmp_path / "m"), sc.lambertian.from_color(sc.color(.1, .2, .3)), scale=2.0, offset=sc.point3(1, 0, 0))
    assert len(m.v0) == 5 and m.num_degenerate == 1
    assert m.v0[0].tolist() == [1, 0, 0] and m.v1[0].tolist() == [3, 0, 0] and m.v2[0].tolist() == [3, 2, 0]
    extra = sc.triangle(sc.point3(0, 0, 5), sc.point3(1, 0, 5), sc.point3(0, 1, 5), sc.metal(sc.color(.5, .5, .5), 0.1))
    a = SC.compile_scene(sc.hittable_list([m, extra]))
    b = SC.compile_scene(sc.hittable_list(list(m.triangles) + [extra]))
    assert a["triangle_v0"].shape == (6, 3) and a["triangle_material_type"].tolist() == [0, 0, 0, 0, 0, 1]
    for k in a:
        if k.startswith("triangle_"):
            assert np.array_equal(a[k], b[k]) and a[k].dtype == b[k].dtype, k
    with pytest.raises(FileNotFoundError):
        sc.mesh(str(tmp_path / "nope"), None)


def test_perlin_tables_follow_the_random_stream():
    random.seed(42)
    a = SC.perlin_tables()
    random.seed(42)
    b = SC.perlin_tables()
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert sorted(a["perlin_perm_x"].tolist()) == list(range(256)) and a["perlin_randvec"].shape == (256, 3)
    assert np.abs(a["perlin_randvec"]).max() <= 1.0


@pytest.mark.skipif(not REF, reason="reference not mounted")
@pytest.mark.parametrize("name", ["wavefront_comparison", "cornell_smoke", "cornell_box", "simple_light", "vol2_final_scene"])
def test_against_reference_compiler(name):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import refimport as R
    world, cam = R.capture_scene(name)
    state = random.getstate()
    ref = R.reference_compile(world)
    mine = SC.compile_scene(world)
    mine.update(bvh.compile_bvh(mine))
    for k, v in ref.items():
        assert np.array_equal(np.asarray(v), np.asarray(mine[k])), k
        assert np.asarray(v).dtype == np.asarray(mine[k]).dtype or k == "num_images", k
    random.setstate(state)
    rp = R.reference_perlin_tables()
    random.setstate(state)
    mp = SC.perlin_tables()
    assert all(np.array_equal(rp[k], mp[k]) for k in rp)
    # camera vectors: my stand-in camera class against the reference's camera.initialize()
    rc = R.camera_arrays(cam)
    c2 = sc.camera()
    for a in ("aspect_ratio", "img_width", "samples_per_pixel", "vfov", "defocus_angle", "focus_distance"):
        setattr(c2, a, getattr(cam, a))
    c2.lookfrom, c2.lookat, c2.vup = cam.lookfrom, cam.lookat, cam.vup
    c2.initialize()
    assert c2.img_height == int(rc["img_height"])
    for k, attr in (("cam_center", "center"), ("cam_pixel00", "pixel00_loc"), ("cam_delta_u", "delta_u"),
                    ("cam_delta_v", "delta_v"), ("cam_defocus_disk_u", "defocus_disk_u")):
        v = getattr(c2, attr)
        assert np.array_equal(rc[k], np.array([v.x, v.y, v.z])), k


@pytest.mark.parametrize("name", ["vol2_final_scene", "wavefront_comparison"])
def test_own_scene_scripts_reproduce_the_reference_scenes(name):
    """scenes.vol2_final_scene / wavefront_comparison (own classes, own compiler, native SAH builder -- what
    bench.py renders) against the golden fixture = the arrays the reference's own scene script + compiler made
    for seed 1234: the same primitives with the same materials (as multisets: the reference's bvh_node sorts
    its objects, the container here does not, so indices differ) and bit-identical Perlin tables."""
    import golden_fixtures
    from path_tracer_python_b200 import scenes
    world, cam = getattr(scenes, name)()
    cs = scenes.compile_world(world)
    g, meta, _ = golden_fixtures.load_fixture(name)

    def rows(d, prefix, geo_keys):
        n = len(np.atleast_2d(d[geo_keys[0]]))
        cols = [np.asarray(d[k], np.float64).reshape(n, -1) for k in geo_keys]
        cols += [np.asarray(d[f"{prefix}_{k}"], np.float64).reshape(n, -1) for k in (
            "material_type", "material_albedo", "material_fuzz", "material_ir", "material_emit_color", "texture_type",
            "texture_scale", "texture_color1", "texture_color2", "texture_image_idx", "is_constant_medium", "medium_density",
            "medium_albedo")]
        a = np.concatenate(cols, axis=1)
        return a[np.lexsort(a.T[::-1])]

    assert np.array_equal(rows(cs, "sphere", ["sphere_data"]), rows(g, "sphere", ["sphere_data"]))
    if len(g["quad_Q"]):
        qk = ["quad_Q", "quad_u", "quad_v", "quad_normal", "quad_D", "quad_w"]
        assert np.array_equal(rows(cs, "quad", qk), rows(g, "quad", qk))
    assert len(cs["quad_Q"]) == len(g["quad_Q"]) and len(cs["triangle_v0"]) == len(g["triangle_v0"])
    for k in ("perlin_randvec", "perlin_perm_x", "perlin_perm_y", "perlin_perm_z"):
        assert np.array_equal(cs[k], g[k]), k
    assert int(cs["num_images"]) == int(g["num_images"])
    if int(g["num_images"]):
        assert np.array_equal(cs["image_0"], g["image_0"])
    assert len(cs["bvh_prim_idx"]) == len(g["bvh_prim_idx"])               # 2 * primitives - 1 nodes
    cam.initialize()
    assert (cam.img_width, cam.img_height) == ((1920, 1080) if name == "vol2_final_scene" else (400, 225))
