"""Pins the CPU oracle: bit-for-bit equality with the fixtures produced by executing the
reference's own renderer.py / kernels.py / fields.py under oracle/taichi_shim (gen_golden.py)."""
import numpy as np
import pytest

from conftest import GOLDEN_SCENES
from path_tracer_python_b200 import scenes


@pytest.fixture(scope="module", params=GOLDEN_SCENES)
def case(request, oracle_mod):
    O = oracle_mod
    cs, meta, out = scenes.load_fixture(request.param)
    S = O.OracleScene(cs)
    cam = O.make_camera(meta["camera"], meta["background"], meta["max_depth"])
    return O, S, cam, meta, out


def test_camera_rays(case):
    O, S, cam, meta, out = case
    o, d, _ = O.generate_camera_rays(cam, meta["width"], meta["height"], meta["rng_seed"], 0)
    assert np.array_equal(o, out["ray_o"]) and np.array_equal(d, out["ray_d"])


def test_first_hit_table(case):
    O, S, cam, meta, out = case
    fh = O.first_hit(S, out["ray_o"], out["ray_d"])
    assert np.array_equal(fh["valid"], out["hit_valid"])
    assert np.array_equal(fh["t"], out["hit_t"])
    assert np.array_equal(fh["prim_type"], out["hit_prim_type"])
    assert np.array_equal(fh["prim_idx"], out["hit_prim_idx"])
    assert np.array_equal(fh["p"], out["hit_point"])
    assert np.array_equal(fh["n"], out["hit_normal"])


@pytest.mark.parametrize("mode", ["megakernel", "wavefront"])
def test_accumulation_buffer(case, mode):
    O, S, cam, meta, out = case
    acc, cnt = O.render(S, cam, meta["width"], meta["height"], meta["spp"], seed=meta["rng_seed"], mode=mode)
    assert np.array_equal(acc, out["accum_" + mode])
    assert cnt["paths"] == meta["width"] * meta["height"] * meta["spp"]


def test_sample_and_row_ranges_compose(case):
    """Rendering [0,a)+[a,spp) samples, or two row bands, equals one call (the multi-GPU split)."""
    O, S, cam, meta, out = case
    W, H, spp, seed = meta["width"], meta["height"], meta["spp"], meta["rng_seed"]
    acc = np.zeros((H, W, 3), np.float32)
    O.render(S, cam, W, H, 1, seed=seed, accum=acc)
    O.render(S, cam, W, H, spp - 1, seed=seed, accum=acc, sample_begin=1)
    assert np.array_equal(acc, out["accum_megakernel"])
    acc = np.zeros((H, W, 3), np.float32)
    O.render(S, cam, W, H, spp, seed=seed, accum=acc, y0=0, y1=H // 2)
    O.render(S, cam, W, H, spp, seed=seed, accum=acc, y0=H // 2, y1=H)
    assert np.array_equal(acc, out["accum_megakernel"])


def test_tonemap_matches_reference_formula(oracle_mod):
    rng = np.random.default_rng(0)
    acc = (rng.random((17, 23, 3), dtype=np.float32) * 40 - 2).astype(np.float32)
    from path_tracer_python_b200.preview import buffer_to_image
    for spp in (1, 7, 256):
        # preview.py:117-132
        scale = 1.0 / max(1, spp)
        ref = np.clip(np.sqrt(np.maximum(0, acc * scale)) * 255.999, 0, 255).astype(np.uint8)
        assert np.array_equal(oracle_mod.buffer_to_image(acc, spp), ref)
        assert np.array_equal(buffer_to_image(acc, spp), ref)


def test_contract_math_accuracy(oracle_mod):
    """The contract's elementary functions are accurate stand-ins for libm (a few ulp)."""
    O = oracle_mod
    x = np.linspace(-300, 300, 200001).astype(np.float32)
    assert np.abs(O.math_eval("sin", x) - np.sin(x.astype(np.float64))).max() < 2e-6
    assert np.abs(O.math_eval("cos", x) - np.cos(x.astype(np.float64))).max() < 2e-6
    u = np.linspace(1e-10, 1, 100001).astype(np.float32)
    assert np.abs(O.math_eval("log", u) - np.log(u.astype(np.float64))).max() < 3e-6
    c = np.linspace(-1, 1, 100001).astype(np.float32)
    assert np.abs(O.math_eval("acos", c) - np.arccos(c.astype(np.float64))).max() < 1e-6
    a = np.random.default_rng(1).normal(size=100000).astype(np.float32)
    b = np.random.default_rng(2).normal(size=100000).astype(np.float32)
    assert np.abs(O.math_eval("atan2", a, b) - np.arctan2(a.astype(np.float64), b.astype(np.float64))).max() < 1e-6


def test_rng_stream_uniform(oracle_mod):
    O = oracle_mod
    s = O.rng_seed(1234, 77, 3)
    v, _ = O.rng_draw(s, 200000)
    assert v.min() >= 0.0 and v.max() < 1.0
    assert abs(v.mean() - 0.5) < 3e-3 and abs(v.var() - 1 / 12) < 2e-3
    assert O.rng_seed(1234, 77, 3) != O.rng_seed(1234, 77, 4) != O.rng_seed(1234, 78, 3)
