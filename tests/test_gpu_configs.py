"""BASELINE.json configs 3 (1 M-triangle mesh) and 4 (Cornell + medium + glass + SSS-as-Lambert)."""
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from path_tracer_python_b200.engine import Engine
    return Engine(0)


def test_config3_million_triangles_first_hit_bit_exact(engine, oracle_mod):
    """~1 M triangles: native SAH build, then first-hit parity (bit-exact t / point / normal, ids equal
    up to exact-t ties between neighbouring triangles) on primary and on incoherent bounce rays."""
    from path_tracer_python_b200 import scenes
    O = oracle_mod
    t0 = time.time()
    cs, meta = scenes.heightfield_scene(img_width=480)
    build_s = time.time() - t0
    assert len(cs["triangle_v0"]) == 1_000_000 and len(cs["bvh_prim_idx"]) == 2 * 1_000_002 - 1
    assert build_s < 60
    info = engine.set_scene(cs)
    assert info.max_depth_tree <= 64
    engine.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    W, H = meta["width"], meta["height"]
    o, d = engine.generate_camera_rays(W, H, 5, 0)
    # second-generation rays: from the first hit points into random directions
    t, prim, p, n = engine.first_hit(o, d)
    g = torch.Generator(device="cpu").manual_seed(1)
    rd = torch.randn((o.shape[0], 3), generator=g).to(o.device)
    rays = [(o, d), (p + 1e-3 * n, rd)]
    S = O.OracleScene(cs)
    for ro, rdir in rays:
        sel = torch.arange(0, ro.shape[0], 7, device=ro.device)       # oracle on a 1/7 subset (~18 k rays)
        ro, rdir = ro[sel].contiguous(), rdir[sel].contiguous()
        gt, gprim, gp, gn = (x.cpu().numpy() for x in engine.first_hit(ro, rdir))
        ref = O.first_hit(S, ro.cpu().numpy(), rdir.cpu().numpy())
        assert np.array_equal(gprim >= 0, ref["valid"] == 1)
        assert np.array_equal(gt, ref["t"])
        rp = (ref["prim_type"].astype(np.int64) << 28) | ref["prim_idx"]
        hit = ref["valid"] == 1
        same = gprim[hit] == rp[hit]
        assert same.mean() > 0.995
        m = hit.copy(); m[hit] = same
        assert np.array_equal(gp[m], ref["p"][m]) and np.array_equal(gn[m], ref["n"][m])
        assert hit.mean() > 0.3


def test_config3_paths_megakernel_equals_oracle_on_a_window(engine, oracle_mod):
    from path_tracer_python_b200 import scenes
    O = oracle_mod
    cs, meta = scenes.heightfield_scene(img_width=96)
    engine.set_scene(cs)
    engine.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    W, H = meta["width"], meta["height"]
    S = O.OracleScene(cs); cam = O.make_camera(meta["camera"], meta["background"], meta["max_depth"])
    for mode, fn in (("megakernel", engine.render_megakernel), ("wavefront", engine.render_wavefront)):
        acc = engine.new_accum(W, H); fn(acc, 0, 4, 21)
        ref, _ = O.render(S, cam, W, H, 4, seed=21, mode=mode)
        assert np.array_equal(acc.cpu().numpy(), ref), mode


def _config4(width, spp):
    from path_tracer_python_b200 import scenes
    from path_tracer_python_b200.renderer import TaichiRenderer
    world, cam = scenes.cornell_config4(width, spp)
    return world, cam, TaichiRenderer


def test_config4_cornell_bit_exact_and_converges(oracle_mod, tmp_path):
    """Reduced-size config 4 through the public API: bit-exact vs the oracle at 96x96x8, and
    convergence at 128x128: two independent 512-spp renders agree within Monte-Carlo error
    (per-pixel RMSE <= 3x the standard error estimated from the two) and in mean luminance (0.5 %)."""
    O = oracle_mod
    world, cam, R = _config4(96, 8)
    r = R(world, cam, str(tmp_path / "c4.png"), seed=3, quiet=True)
    r.background_color = (0, 0, 0)
    r.max_depth = 50
    assert r.info.has_media == 1 and r.num_spheres == 2
    assert r.compiled["sphere_material_type"].tolist() == [2, 0]            # glass, SSS -> grey Lambert
    r.render(enable_preview=False)
    S = O.OracleScene(r.compiled)
    ocam = O.make_camera(r._camera_dict(), (0, 0, 0), 50)
    ref, _ = O.render(S, ocam, 96, 96, 8, seed=3)
    assert np.array_equal(r.accum_numpy(), ref)

    imgs = []
    for seed in (101, 202):
        world, cam, R = _config4(128, 512)
        rr = R(world, cam, str(tmp_path / f"c4_{seed}.png"), seed=seed, quiet=True)
        rr.background_color = (0, 0, 0)
        rr.render_wavefront(enable_preview=False)
        imgs.append(rr.accum_numpy() / 512.0)
    a, b = imgs
    assert np.isfinite(a).all() and np.isfinite(b).all()
    lum = lambda x: (0.2126 * x[..., 0] + 0.7152 * x[..., 1] + 0.0722 * x[..., 2])
    assert abs(lum(a).mean() - lum(b).mean()) / lum(a).mean() < 0.01
    # the difference of two independent unbiased estimates has variance 2*sigma^2/spp per pixel;
    # a blurred version must shrink like noise does (no structured bias between seeds)
    diff = lum(a) - lum(b)
    blk = diff.reshape(16, 8, 16, 8).mean(axis=(1, 3))
    assert blk.std() < 0.35 * diff.std() + 1e-6
    assert 0.05 < lum(a).mean() < 5.0
