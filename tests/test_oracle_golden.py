"""Pins the CPU oracle: bit-for-bit equality with the fixtures produced by executing the
reference's own renderer.py / kernels.py / fields.py under oracle/taichi_shim (gen_golden.py)."""
import numpy as np
import pytest

from conftest import GOLDEN_SCENES
from path_tracer_python_b200 import scenes
import golden_fixtures


@pytest.fixture(scope="module", params=GOLDEN_SCENES)
def case(request, oracle_mod):
    O = oracle_mod
    cs, meta, out = golden_fixtures.load_fixture(request.param)
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


def test_rng_no_two_paths_share_a_stream(oracle_mod):
    """Config 5 has 3840*2160*1024 = 8.5e9 paths > 2^32: the stream identity must be wider than 32
    bits.  (start, stride) is a bijective function of (pixel, sample) for one seed; checked here on
    1.7e7 keys of config 5's domain: no duplicate stream state, no duplicate 2-draw (64-bit) prefix."""
    O = oracle_mod

    def n_distinct(a):                                                     # (np.unique is slow on 1e7 uint64)
        b = np.sort(np.ascontiguousarray(a).ravel())
        return int(1 + np.count_nonzero(b[1:] != b[:-1]))

    npix = 3840 * 2160
    g = np.random.default_rng(5)
    groups = [(np.arange(npix, dtype=np.uint32), np.zeros(npix, np.uint32)),                     # one sample of every pixel
              (g.integers(0, npix, npix, dtype=np.uint32), g.integers(1, 1024, npix, dtype=np.uint32)),
              (np.full(1 << 20, 12345, np.uint32), np.arange(1024, 1024 + (1 << 20), dtype=np.uint32))]   # one pixel, 2^20 samples
    px, sm = (np.concatenate(c) for c in zip(*groups))
    keys = (px.astype(np.uint64) << np.uint64(32)) | sm.astype(np.uint64)
    st = O.rng_seed_states(1234, px, sm)
    assert keys.size > 1.7e7 and n_distinct(st) == n_distinct(keys)   # distinct keys <=> distinct streams
    assert np.all((st >> np.uint64(32)) & np.uint64(1) == 1)                 # strides are odd: full 2^32 period
    # every pixel of one sample starts somewhere else; every sample of one pixel has its own stride
    assert n_distinct(st[:npix] & np.uint64(0xFFFFFFFF)) > 0.99 * npix
    one_pixel = st[-(1 << 20):]
    assert n_distinct(one_pixel >> np.uint64(32)) == one_pixel.size
    d = O.rng_first_draws(st, 2)
    prefix = np.ascontiguousarray(d).view(np.uint64).ravel()                 # the first two raw draws of every stream
    assert n_distinct(prefix) == n_distinct(keys)
    # neighbouring streams are uncorrelated draw by draw
    u = d[:npix].astype(np.float64) / 2.0 ** 32
    assert abs(np.corrcoef(u[:-1, 0], u[1:, 0])[0, 1]) < 2e-3 and abs(np.corrcoef(u[:, 0], u[:, 1])[0, 1]) < 2e-3
