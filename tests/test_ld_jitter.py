"""jitter="sobol": low-discrepancy sub-pixel offsets (docs/__optimizations.md:7-10) -- the
sample's point of a per-pixel XOR-scrambled Sobol' (0,2)-sequence instead of two random draws (kernels.py:183-184).
Same expected image (the points are uniform), better-stratified pixel footprints.  No reference code to compare with:
CPU tests pin the net property and the expectation, the GPU test pins CUDA against the oracle bit for bit."""
import numpy as np
import pytest

import golden_fixtures

lum = lambda x: 0.2126 * x[..., 0] + 0.7152 * x[..., 1] + 0.0722 * x[..., 2]


def test_points_form_a_02_net_per_pixel(oracle_mod):
    O = oracle_mod
    for seed, pixel in ((1234, 0), (1234, 2073599), (7, 99)):
        p = O.ld_jitter_points(seed, pixel, 256).astype(np.float64)
        assert p.min() >= 0.0 and p.max() < 1.0
        # every block of 2^k consecutive samples (aligned) puts exactly one point in each cell of every 2^a x 2^(k-a) grid
        for k in (2, 4, 6, 8):
            n = 1 << k
            for start in range(0, 256, n):
                q = p[start:start + n]
                for a in range(k + 1):
                    cells = (np.floor(q[:, 0] * (1 << a)).astype(int) << (k - a)) | np.floor(q[:, 1] * (1 << (k - a))).astype(int)
                    assert len(set(cells.tolist())) == n, (seed, pixel, k, start, a)
    a, b = O.ld_jitter_points(1234, 5, 64), O.ld_jitter_points(1234, 6, 64)
    assert not np.allclose(a, b)                                         # neighbouring pixels are scrambled differently


def test_same_expectation_less_noise_where_the_pixel_footprint_matters(oracle_mod):
    """Depth 1 (visibility + emission + background only): all the noise is anti-aliasing noise."""
    O = oracle_mod
    cs, meta, _ = golden_fixtures.load_fixture("wavefront_comparison")
    from path_tracer_python_b200.scenes import _CAMERAS, camera_vectors
    cam, H = camera_vectors(img_width=64, aspect_ratio=16 / 9, **_CAMERAS["wavefront_comparison"])
    S = O.OracleScene(cs)
    imgs = {}
    for jitter in ("random", "sobol"):
        oc = O.make_camera(cam, (0.7, 0.8, 1.0), 1, jitter)
        imgs[jitter] = [O.render(S, oc, 64, H, 64, seed=s, mode="wavefront")[0].astype(np.float64) / 64 for s in (1, 2, 3, 4)]
    mean = {k: np.mean(v, axis=0) for k, v in imgs.items()}
    noise = {k: np.sqrt(np.mean([(lum(x) - lum(mean[k])) ** 2 for x in v])) for k, v in imgs.items()}
    assert abs(lum(mean["sobol"]).mean() / lum(mean["random"]).mean() - 1) < 0.003
    assert np.sqrt(np.mean((lum(mean["sobol"]) - lum(mean["random"])) ** 2)) < 2.0 * noise["random"]
    assert noise["sobol"] < 0.6 * noise["random"]
    # full paths: still the same image
    full = {}
    for jitter in ("random", "sobol"):
        oc = O.make_camera(cam, (0.7, 0.8, 1.0), 50, jitter)
        full[jitter] = O.render(S, oc, 64, H, 256, seed=9, mode="megakernel")[0].astype(np.float64) / 256
    assert abs(lum(full["sobol"]).mean() / lum(full["random"]).mean() - 1) < 0.01


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["wavefront_comparison", "feature_zoo"])
def test_gpu_sobol_jitter_bit_exact_vs_oracle(oracle_mod, name):
    from path_tracer_python_b200.engine import Engine
    O = oracle_mod
    cs, meta, _ = golden_fixtures.load_fixture(name)
    e = Engine(0)
    e.set_scene(cs)
    e.set_camera(meta["camera"], meta["background"], meta["max_depth"], jitter="sobol")
    W, H = meta["width"], meta["height"]
    S = O.OracleScene(cs)
    oc = O.make_camera(meta["camera"], meta["background"], meta["max_depth"], "sobol")
    o, d = e.generate_camera_rays(W, H, 77, 3)
    ro, rd, _ = O.generate_camera_rays(oc, W, H, 77, 3)
    assert np.array_equal(o.cpu().numpy(), ro) and np.array_equal(d.cpu().numpy(), rd)
    for mode, fn in (("megakernel", e.render_megakernel), ("wavefront", e.render_wavefront)):
        acc = e.new_accum(W, H)
        fn(acc, 2, 5, 77)                                   # samples 2..6: the sequence index is the global sample index
        ref, _ = O.render(S, oc, W, H, 5, seed=77, mode=mode, sample_begin=2)
        assert np.array_equal(acc.cpu().numpy(), ref), mode
    with pytest.raises(Exception, match="jitter"):
        e.set_camera(meta["camera"], meta["background"], meta["max_depth"], jitter="halton")
