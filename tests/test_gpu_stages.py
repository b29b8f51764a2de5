"""The wavefront pipeline one kernel at a time through the stage-level C ABI (pt_wf_generate / pt_wf_extend /
pt_wf_shade / pt_wf_compact / pt_wf_resolve), against the stage outputs the reference's own kernels produced
(tests/golden: ray_origins / ray_directions after generate_camera_rays, hit_* after intersect_rays) and, driven
by the reference's host loop (renderer.py:305-334), against its final accumulation buffer."""
import numpy as np
import pytest
import torch

import golden_fixtures
from conftest import GOLDEN_SCENES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from path_tracer_python_b200.engine import Engine
    return Engine(0)


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_generate_and_extend_match_the_reference_kernels(engine, name):
    cs, meta, out = golden_fixtures.load_fixture(name)
    engine.set_scene(cs)
    engine.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    W, H, seed = meta["width"], meta["height"], meta["rng_seed"]
    n = W * H
    engine.wf_begin(W, H, 0, 1, seed)
    engine.wf_generate()
    assert engine.wf_compact(0) == n
    rays = {k: v.cpu().numpy() for k, v in engine.wf_read_rays(0, n).items()}
    order = np.argsort(rays["pixel"], kind="stable")
    assert np.array_equal(rays["pixel"][order], np.arange(n))                   # one camera ray per pixel
    assert np.array_equal(rays["o"][order], out["ray_o"]) and np.array_equal(rays["d"][order], out["ray_d"])
    assert np.all(rays["throughput"] == 1.0) and np.all(rays["depth"] == 0)
    engine.wf_extend(0)
    hits = {k: v.cpu().numpy() for k, v in engine.wf_read_hits(n).items()}
    t, prim = hits["t"][order], hits["prim"][order]
    valid = out["hit_valid"] == 1
    assert np.array_equal(prim >= 0, valid)
    # a medium boundary is recorded with its entry t; the reference's intersect_rays records the same closest hit
    assert np.array_equal(t[valid], out["hit_t"][valid])
    ref_prim = (out["hit_prim_type"].astype(np.int64) << 28) | out["hit_prim_idx"]
    same = prim[valid] == ref_prim[valid]
    assert same.all()                                                            # exact-t ties included (reference visiting order)


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_host_loop_over_the_stages_reproduces_the_reference_image(engine, name):
    """for sample: generate; for bounce < max_depth: (stop when the queue is empty) extend, shade -- the reference's
    render_wavefront loop, one stage call per kernel -- equals its accumulation buffer bit for bit."""
    cs, meta, out = golden_fixtures.load_fixture(name)
    engine.set_scene(cs)
    engine.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    W, H, spp, seed, max_depth = meta["width"], meta["height"], meta["spp"], meta["rng_seed"], meta["max_depth"]
    acc = engine.new_accum(W, H)
    waves = []
    for s in range(spp):
        engine.wf_begin(W, H, s, 1, seed)
        engine.wf_generate()
        for bounce in range(max_depth):
            live = engine.wf_compact(bounce)
            if live == 0:
                break
            waves.append(live)
            engine.wf_extend(bounce)
            engine.wf_shade(bounce)
        engine.wf_resolve(acc)
    assert np.array_equal(acc.cpu().numpy(), out["accum_wavefront"])
    assert waves[0] == W * H and min(waves) < W * H
    # the same frame as ONE batch of all samples (what pt_render_wavefront does) through the same stage calls
    acc2 = engine.new_accum(W, H)
    engine.wf_begin(W, H, 0, spp, seed)
    engine.wf_generate()
    for bounce in range(max_depth):
        if engine.wf_compact(bounce) == 0:
            break
        engine.wf_extend(bounce)
        engine.wf_shade(bounce)
    engine.wf_resolve(acc2)
    assert torch.equal(acc, acc2)


def test_stage_calls_need_a_bound_batch():
    from path_tracer_python_b200 import _lib
    from path_tracer_python_b200.engine import Engine
    e = Engine(0)
    with pytest.raises(_lib.B200PTError, match="pt_wf_begin"):
        e.wf_generate()
