"""`full` feature mode: motion blur (moving spheres) and the two subsurface materials.

The reference's Taichi path drops both (SURVEY.md "three facts" #2); they exist only in its
pure-Python tracer.  Parity is therefore two-step:
  1. CPU: the oracle's full mode against a radiance image rendered by the reference's own
     `camera.ray_color` loop (fixture written by oracle/gen_golden_full.py) -- statistically;
  2. GPU: the CUDA full mode against the oracle's full mode -- bit for bit, both architectures.
"""
import numpy as np
import pytest

from path_tracer_python_b200 import scenes
import golden_fixtures


def _load():
    cs, meta, out = golden_fixtures.load_fixture("full_features")
    return cs, meta, out["radiance_pure_python"]


def test_fixture_is_a_full_feature_scene():
    cs, meta, ref = _load()
    assert int(cs["full_features"]) == 1
    assert cs["sphere_material_type"].tolist() == [0, 0, 6, 5]                 # Lambert x2, SSS volumetric, SSS simple
    assert np.allclose(cs["sphere_motion"][1], [0.0, 0.6, 0.0]) and not cs["sphere_motion"][[0, 2, 3]].any()
    # the moving sphere's leaf box is the union of both end positions (core/sphere.py:27-32)
    leaf = np.where((cs["bvh_prim_type"] == 0) & (cs["bvh_prim_idx"] == 1))[0][0]
    assert np.allclose(cs["bvh_bbox_min"][leaf], [-2.1, 0.0, -0.5]) and np.allclose(cs["bvh_bbox_max"][leaf], [-1.1, 1.6, 0.5])
    assert ref.shape == (meta["height"], meta["width"], 3)


@pytest.mark.parametrize("mode", ["megakernel", "wavefront"])
def test_oracle_full_mode_matches_pure_python_tracer(oracle_mod, mode):
    O = oracle_mod
    cs, meta, ref = _load()
    W, H = meta["width"], meta["height"]
    S = O.OracleScene(cs)
    cam = O.make_camera(meta["camera"], meta["background"], meta["max_depth"])
    spp = 2048
    acc, _ = O.render(S, cam, W, H, spp, seed=11, mode=mode)
    img = acc / spp
    assert abs(img.mean() / ref.mean() - 1.0) < 0.01                         # mean radiance within 1 %
    a = img.reshape(4, 5, 8, 5, 3).mean(axis=(1, 3, 4))
    b = ref.reshape(4, 5, 8, 5, 3).mean(axis=(1, 3, 4))
    assert np.abs(a / b - 1.0).max() < 0.06                                  # every 5x5 block within MC error
    # the features are really on: compat mode (static sphere, grey Lambert SSS) gives a different image
    compat = dict(cs, full_features=0)
    acc2, _ = O.render(O.OracleScene(compat), cam, W, H, 256, seed=11, mode=mode)
    c = (acc2 / 256).reshape(4, 5, 8, 5, 3).mean(axis=(1, 3, 4))
    assert np.abs(c / b - 1.0).max() > 0.10


def test_contract_exp_accuracy(oracle_mod):
    x = np.linspace(-80, 5, 200001).astype(np.float32)
    got = oracle_mod.math_eval("exp", x).astype(np.float64)
    assert np.abs(got / np.exp(x.astype(np.float64)) - 1).max() < 3e-7


@pytest.mark.gpu
def test_gpu_full_mode_bit_exact_vs_oracle(oracle_mod):
    import torch
    from path_tracer_python_b200.engine import Engine
    O = oracle_mod
    cs, meta, _ = _load()
    W, H = meta["width"], meta["height"]
    e = Engine(0)
    info = e.set_scene(cs)
    assert info.full_features == 1
    e.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    S = O.OracleScene(cs)
    cam = O.make_camera(meta["camera"], meta["background"], meta["max_depth"])
    x = torch.linspace(-100, 5, 300001)
    assert np.array_equal(e.math_eval(7, x).cpu().numpy().view(np.uint32), O.math_eval("exp", x.numpy()).view(np.uint32))
    for mode, fn in (("megakernel", e.render_megakernel), ("wavefront", e.render_wavefront)):
        acc = e.new_accum(W, H)
        fn(acc, 0, 64, 5)
        ref, _ = O.render(S, cam, W, H, 64, seed=5, mode=mode)
        got = acc.cpu().numpy()
        bad = np.abs(got - ref).max(axis=2) > 0
        assert bad.sum() == 0, f"{mode}: {bad.sum()} pixels differ (max {np.abs(got - ref).max()})"


@pytest.mark.gpu
def test_renderer_features_kwarg(tmp_path):
    """Through the public API with this repo's scene classes: features='full' vs default."""
    from path_tracer_python_b200 import scene as sc
    from path_tracer_python_b200.renderer import TaichiRenderer
    w = sc.hittable_list()
    w.add(sc.Sphere.stationary(sc.point3(0, -1000, 0), 1000, sc.lambertian.from_color(sc.color(.5, .5, .5))))
    w.add(sc.Sphere.moving(sc.point3(-1, .5, 0), sc.point3(-1, 1.2, 0), .5, sc.lambertian.from_color(sc.color(.7, .3, .1))))
    w.add(sc.Sphere.stationary(sc.point3(1, .5, 0), .5, sc.subsurface_volumetric(sc.color(.2, .5, .2), .08, .8, .7)))
    imgs = {}
    for feat in ("taichi_compat", "full"):
        cam = sc.camera()
        cam.aspect_ratio, cam.img_width, cam.samples_per_pixel = 2.0, 64, 64
        cam.vfov, cam.lookfrom, cam.lookat = 30, sc.point3(0, 1.5, -6), sc.point3(0, .6, 0)
        r = TaichiRenderer(w, cam, str(tmp_path / f"{feat}.png"), seed=1, quiet=True, features=feat)
        r.render(enable_preview=False)
        imgs[feat] = r.accum_numpy() / 64
        assert r.info.full_features == (1 if feat == "full" else 0)
    assert np.isfinite(imgs["full"]).all()
    assert np.abs(imgs["full"] - imgs["taichi_compat"]).mean() > 1e-3       # motion blur + SSS change the picture
