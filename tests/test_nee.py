"""features="nee": next-event estimation at Lambertian hits (the reference's docs claim it,
docs/optimizations.md:53, and list it as missing, docs/__optimizations.md:7-10 -- there is no reference code).
  CPU: the oracle's NEE estimator has the same expectation as the oracle's (= the reference's) estimator, on scenes with
       glass, constant media and several emitter types, and less noise; the packer's light table.
  GPU: the CUDA NEE kernels against the oracle's NEE restatement, bit for bit, both architectures."""
import ctypes as C

import numpy as np
import pytest

import golden_fixtures

lum = lambda x: 0.2126 * x[..., 0] + 0.7152 * x[..., 1] + 0.0722 * x[..., 2]


def _cornell(features, width=48):
    from path_tracer_python_b200 import bvh, scene_compiler as SC, scenes
    world, cam = scenes.cornell_config4(width, 1)
    cs = SC.compile_scene(world, features)
    cs.update(bvh.compile_bvh(cs))
    cs.update(SC.perlin_tables())
    return cs, scenes.camera_dict(cam), (0.0, 0.0, 0.0), 50, width, cam.img_height


def _zoo(nee):
    cs, meta, _ = golden_fixtures.load_fixture("feature_zoo")
    if nee:
        cs = dict(cs, full_features=2, sphere_motion=np.zeros((len(cs["sphere_data"]), 3), np.float32))
    return cs, meta["camera"], meta["background"], meta["max_depth"], meta["width"], meta["height"]


def _two_lights():
    """A sphere light and a triangle light over a checkered floor, a glass ball and a fog box in between.  (The fog box
    floats: a boundary face flush with the floor would put two different materials at exactly the same t, and which of
    them a traversal meets first is the one thing the CUDA kernels and the reference order do not share -- the
    exact-tie excuse of SURVEY.md section 8a note vi, which then shows in the radiance, with or without NEE.)"""
    from path_tracer_python_b200 import scene as sc, scenes
    P, V, Col = sc.point3, sc.vec3, sc.color
    w = sc.hittable_list()
    w.add(sc.quad(P(-6, 0, -6), V(12, 0, 0), V(0, 0, 12), sc.lambertian.from_texture(
        sc.checker_texture.from_colors(1.0, Col(0.8, 0.8, 0.8), Col(0.3, 0.5, 0.3)))))
    w.add(sc.Sphere.stationary(P(-2.0, 3.0, 0.0), 0.5, sc.diffuse_light.from_color(Col(12, 10, 8))))
    w.add(sc.triangle(P(1.5, 3.5, -1.0), P(3.0, 3.5, -1.0), P(2.2, 3.5, 0.6), sc.diffuse_light.from_color(Col(6, 8, 12))))
    w.add(sc.Sphere.stationary(P(0.0, 0.8, 0.0), 0.8, sc.dielectric(1.5)))
    w.add(sc.Sphere.stationary(P(2.0, 0.6, 1.5), 0.6, sc.lambertian.from_color(Col(0.7, 0.3, 0.2))))
    w.add(sc.constant_medium.from_color(sc.box(P(-3.5, 0.25, 1.0), P(-1.5, 1.75, 2.5), sc.lambertian.from_color(Col(1, 1, 1))),
                                        Col(0.9, 0.9, 0.9), 0.8))
    cam = sc.camera()
    cam.aspect_ratio, cam.img_width, cam.vfov = 4 / 3, 40, 45
    cam.lookfrom, cam.lookat, cam.vup, cam.defocus_angle = P(0, 4.5, 9), P(0, 1, 0), V(0, 1, 0), 0
    return w, cam


def _lights_scene(features):
    from path_tracer_python_b200 import bvh, scene_compiler as SC, scenes
    world, cam = _two_lights()
    cs = SC.compile_scene(world, features)
    cs.update(bvh.compile_bvh(cs))
    cs.update(SC.perlin_tables())
    return cs, scenes.camera_dict(cam), (0.02, 0.02, 0.03), 30, cam.img_width, cam.img_height


SCENES = {"cornell": lambda f: _cornell(f), "zoo": lambda f: _zoo(f == "nee"), "two_lights": _lights_scene}


@pytest.mark.parametrize("name", ["cornell", "zoo", "two_lights"])
@pytest.mark.parametrize("mode", ["megakernel", "wavefront"])
def test_oracle_nee_has_the_reference_estimators_expectation(oracle_mod, name, mode):
    O = oracle_mod
    out = {}
    for feat, spp in (("taichi_compat", 1536), ("nee", 768)):
        cs, cam, bg, depth, W, H = SCENES[name](feat)
        assert int(cs.get("full_features", 0)) == (2 if feat == "nee" else 0)
        S = O.OracleScene(cs)
        oc = O.make_camera(cam, bg, depth)
        a = O.render(S, oc, W, H, spp, seed=11, mode=mode)[0].astype(np.float64) / spp
        b = O.render(S, oc, W, H, spp, seed=22, mode=mode)[0].astype(np.float64) / spp
        out[feat] = (a, b, spp)
    (sa, sb, s_spp), (na, nb, n_spp) = out["taichi_compat"], out["nee"]
    std, nee = 0.5 * (sa + sb), 0.5 * (na + nb)
    # mean luminance: within 3 standard errors of the difference (each estimated from its two seeds), and within 1.5 %
    se_std = abs(lum(sa).mean() - lum(sb).mean()) / 2 + 1e-4 * lum(std).mean()
    se_nee = abs(lum(na).mean() - lum(nb).mean()) / 2 + 1e-4 * lum(std).mean()
    assert abs(lum(nee).mean() - lum(std).mean()) < 3 * np.hypot(se_std, se_nee) + 0.005 * lum(std).mean()
    assert abs(lum(nee).mean() / lum(std).mean() - 1) < 0.015
    # 4x4 blocks: the difference of the two estimators is noise (no structured bias)
    Hb, Wb = (lum(std).shape[0] // 4) * 4, (lum(std).shape[1] // 4) * 4
    blk = lambda x: lum(x)[:Hb, :Wb].reshape(Hb // 4, 4, Wb // 4, 4).mean(axis=(1, 3))
    noise = np.sqrt((blk(sa) - blk(sb)) ** 2 / 4 + (blk(na) - blk(nb)) ** 2 / 4).mean()
    assert np.sqrt(np.mean((blk(nee) - blk(std)) ** 2)) < 3.0 * noise + 1e-3
    # less noise per sample: NEE with half the samples is not noisier than the plain estimator
    rms = lambda x, y: np.sqrt(np.mean((lum(x) - lum(y)) ** 2))
    assert rms(na, nb) * np.sqrt(n_spp) < rms(sa, sb) * np.sqrt(s_spp)


def test_light_table_of_the_packed_scene():
    from path_tracer_python_b200 import _lib
    cs, *_ = _lights_scene("nee")
    hs = _lib.HostScene(cs)
    info = _lib.PtPackedInfo()
    L = _lib.lib()
    _lib.check(L.pt_pack_scene_info(C.byref(hs.c), C.byref(info)), "info")
    assert info.num_lights == 2 and info.full_features == 2
    blob = np.zeros(int(info.total_bytes), np.uint8)
    _lib.check(L.pt_pack_scene(C.byref(hs.c), C.byref(info), blob.ctypes.data_as(C.c_void_p)), "pack")
    rec = blob[int(info.off_lights): int(info.off_lights) + 64].view(np.float32).reshape(2, 8)
    refs = ~rec[:, 0].copy().view(np.int32)
    assert (refs & 3).tolist() == [0, 1] and ((refs >> 2) & 7).tolist() == [1, 1]      # a sphere and a triangle, class "emissive"
    assert np.isclose(rec[0, 1], 4 * np.pi * 0.25) and np.isclose(rec[1, 1], 0.5 * 1.5 * 1.6)
    assert np.allclose(rec[0, 2:5], [12, 10, 8]) and np.allclose(rec[1, 2:5], [6, 8, 12])
    assert np.isclose(info.nee_r2, 0.25 * (rec[0, 1] + rec[1, 1]) / 2)
    cs0, *_ = _lights_scene("taichi_compat")
    hs0 = _lib.HostScene(cs0)
    _lib.check(L.pt_pack_scene_info(C.byref(hs0.c), C.byref(info)), "info")
    assert info.num_lights == 0 and info.full_features == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cornell", "zoo", "two_lights"])
def test_gpu_nee_bit_exact_vs_oracle(oracle_mod, name):
    import torch
    from path_tracer_python_b200.engine import Engine
    O = oracle_mod
    cs, cam, bg, depth, W, H = SCENES[name]("nee")
    e = Engine(0)
    info = e.set_scene(cs)
    assert info.num_lights >= 1
    e.set_camera(cam, bg, depth)
    S = O.OracleScene(cs)
    oc = O.make_camera(cam, bg, depth)
    for mode, fn in (("megakernel", e.render_megakernel), ("wavefront", e.render_wavefront)):
        acc = e.new_accum(W, H)
        fn(acc, 0, 6, 5)
        ref, _ = O.render(S, oc, W, H, 6, seed=5, mode=mode)
        got = acc.cpu().numpy()
        assert np.array_equal(got, ref), f"{name} {mode}: {np.count_nonzero(got != ref)} of {got.size} values differ, max {np.abs(got - ref).max()}"
    # batch size / pipelines do not matter
    for flight in (1, 4):
        e.wavefront_samples_in_flight = flight
        acc = e.new_accum(W, H)
        e.render_wavefront(acc, 0, 6, 5)
        assert np.array_equal(acc.cpu().numpy(), O.render(S, oc, W, H, 6, seed=5, mode="wavefront")[0])
    # with the flag off the feature-mode kernels reproduce the plain estimator
    cs0, *_ = SCENES[name]("taichi_compat")
    e.wavefront_samples_in_flight = None
    e.set_scene(cs0)
    acc = e.new_accum(W, H)
    e.render_wavefront(acc, 0, 6, 5)
    assert np.array_equal(acc.cpu().numpy(), O.render(O.OracleScene(cs0), oc, W, H, 6, seed=5, mode="wavefront")[0])


@pytest.mark.gpu
def test_renderer_nee_converges_to_the_plain_image_with_less_noise(tmp_path):
    """Through the public class: TaichiRenderer(..., features="nee") on the Cornell config at 128x128."""
    from path_tracer_python_b200 import scenes
    from path_tracer_python_b200.renderer import TaichiRenderer
    imgs = {}
    for feat, spp in (("taichi_compat", 2048), ("nee", 1024)):
        for seed in (1, 2):
            world, cam = scenes.cornell_config4(128, spp)
            r = TaichiRenderer(world, cam, str(tmp_path / f"{feat}{seed}.png"), seed=seed, quiet=True, features=feat)
            r.background_color = (0, 0, 0)
            r.render_wavefront(enable_preview=False)
            imgs[feat, seed] = r.accum_numpy().astype(np.float64) / spp
    std, nee = 0.5 * (imgs["taichi_compat", 1] + imgs["taichi_compat", 2]), 0.5 * (imgs["nee", 1] + imgs["nee", 2])
    assert abs(lum(nee).mean() / lum(std).mean() - 1) < 0.005
    rms = lambda x, y: np.sqrt(np.mean((lum(x) - lum(y)) ** 2))
    noise_std, noise_nee = rms(imgs["taichi_compat", 1], imgs["taichi_compat", 2]), rms(imgs["nee", 1], imgs["nee", 2])
    print(f"per-pixel noise (two seeds): plain 2048 spp {noise_std:.5f}, nee 1024 spp {noise_nee:.5f}; "
          f"variance per sample {((noise_std ** 2 * 2048) / (noise_nee ** 2 * 1024)):.2f}x lower with nee")
    assert noise_nee * np.sqrt(1024) < noise_std * np.sqrt(2048)
    assert rms(nee, std) < 3 * np.hypot(noise_std, noise_nee) / np.sqrt(2)
