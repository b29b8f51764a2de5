"""The device implementation of the PT-F32 elementary functions / random stream == the oracle's
independent implementation of the same specification, bit for bit, over dense input sweeps."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from path_tracer_python_b200.engine import Engine
    return Engine(0)


def _sweep(lo, hi, n, seed):
    rng = np.random.default_rng(seed)
    x = np.concatenate([np.linspace(lo, hi, n), rng.uniform(lo, hi, n)]).astype(np.float32)
    return x


@pytest.mark.parametrize("fn,idx,lo,hi", [("sin", 0, -700.0, 700.0), ("cos", 1, -700.0, 700.0), ("sin", 0, 0.0, 6.2832),
                                          ("cos", 1, 0.0, 6.2832), ("log", 2, 1e-10, 1.0), ("acos", 3, -1.0001, 1.0001),
                                          ("pow5", 5, 0.0, 1.0)])
def test_unary_bit_exact(engine, oracle_mod, fn, idx, lo, hi):
    x = _sweep(lo, hi, 400000, idx)
    x = np.concatenate([x, np.float32([0.0, -0.0, 1.0, -1.0, 1e-38, 1e-45, 3e7, np.inf])])
    got = engine.math_eval(idx, torch.from_numpy(x)).cpu().numpy()
    ref = oracle_mod.math_eval(fn, x)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_atan2_bit_exact(engine, oracle_mod):
    rng = np.random.default_rng(9)
    a = np.concatenate([rng.normal(size=400000), [0, 0, 1, -1, 0.0, -0.0]]).astype(np.float32)
    b = np.concatenate([rng.normal(size=400000), [0, 1, 0, 0, -1.0, -1.0]]).astype(np.float32)
    got = engine.math_eval(4, torch.from_numpy(a), torch.from_numpy(b)).cpu().numpy()
    ref = oracle_mod.math_eval("atan2", a, b)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_traversal_reciprocal_bit_exact(engine):
    """inv = |d| > 1e-8 ? 1/d : 1e8 (kernels.py:602-607) through the merged-guard fast path ==
    IEEE division, over random bit patterns of every exponent (denormals, huge values, zeros, infs
    included -- those take the library fallback)."""
    rng = np.random.default_rng(11)
    bits = rng.integers(0, 2 ** 32, size=8_000_000, dtype=np.uint64).astype(np.uint32)
    edge = np.float32([0.0, -0.0, 1e-8, -1e-8, 1.0000001e-8, 1e-45, 1e-38, 1.1754944e-38, 8.5e37, 1.7e38, 3.4e38, np.inf, -np.inf,
                       1.0, -1.0, 3.0, 2.0 ** -126, 2.0 ** 126, 2.0 ** 125, 1.9999999 * 2.0 ** 125])
    a = np.concatenate([bits.view(np.float32), edge, np.linspace(-2, 2, 100001, dtype=np.float32)[:-1]])
    a = a[~np.isnan(a)]
    a = a[: len(a) // 2 * 2]
    b = a[::-1].copy()
    got = engine.math_eval(8, torch.from_numpy(a), torch.from_numpy(b)).cpu().numpy()
    x = np.where(np.arange(len(a)) & 1, b, a).astype(np.float32)
    with np.errstate(divide="ignore", over="ignore"):
        ref = np.where(np.abs(x) > np.float32(1e-8), np.float32(1.0) / x, np.float32(1e8)).astype(np.float32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_random_stream_bit_exact(engine, oracle_mod):
    for key in [(1234, 0, 0), (1234, 2073599, 255), (0xFFFFFFFF, 17, 3), (7, 8294399, 1023)]:
        s0 = oracle_mod.rng_seed(*key)                          # stride << 32 | position
        ref, _ = oracle_mod.rng_draw(s0, 4096)
        a = np.zeros(4096, np.uint32)
        a[0], a[1] = s0 & 0xFFFFFFFF, s0 >> 32
        got = engine.math_eval(6, torch.from_numpy(a.view(np.float32))).cpu().numpy()
        assert np.array_equal(got, ref)
