"""The debug-bounds build (libb200pt_dbg.so, -DPT_DEBUG_BOUNDS=1): every node / primitive / material / texel /
queue / bin / slab / stack index the kernels form is checked before use.  compute-sanitizer is not available
on this GPU pool; this is the substitute: the golden suite runs through the checked kernels, must stay
bit-identical to the reference, and must record no violation -- and a deliberately corrupted scene must."""
import numpy as np
import pytest
import torch

import golden_fixtures
from conftest import GOLDEN_SCENES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dbg():
    from path_tracer_python_b200.engine import Engine
    if not Engine.has_numerics("debug"):
        pytest.fail("libb200pt_dbg.so is not built (python -m path_tracer_python_b200.build)")
    e = Engine(0, "debug")
    assert b"debug-bounds" in e.L.pt_version()
    return e


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_golden_suite_through_the_checked_kernels(dbg, name):
    cs, meta, out = golden_fixtures.load_fixture(name)
    dbg.set_scene(cs)
    dbg.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    W, H, spp, seed = meta["width"], meta["height"], meta["spp"], meta["rng_seed"]
    assert dbg.debug_error() == 0
    acc = dbg.new_accum(W, H)
    dbg.render_megakernel(acc, 0, spp, seed)
    assert np.array_equal(acc.cpu().numpy(), out["accum_megakernel"])
    for flight in (1, 2, 32):                       # one / two pipelines, several batch sizes
        dbg.wavefront_samples_in_flight = flight
        acc = dbg.new_accum(W, H)
        dbg.render_wavefront(acc, 0, spp, seed)
        assert np.array_equal(acc.cpu().numpy(), out["accum_wavefront"]), flight
    dbg.wavefront_samples_in_flight = None
    o, d = dbg.generate_camera_rays(W, H, seed, 0)
    dbg.first_hit(o, d)
    code = dbg.debug_error()
    assert code == 0, f"bounds violation: check {code >> 24} at source line {code & 0xffffff}"


def test_bigger_frames_and_every_material(dbg):
    """Config 2 at 480x270 x 32 spp (eight waves + tail, all bins populated) and the 1 M-triangle mesh."""
    from path_tracer_python_b200 import scenes
    world, cam = scenes.vol2_final_scene(480, 16 / 9, 32)
    cs = scenes.compile_world(world)
    dbg.set_scene(cs)
    dbg.set_camera(scenes.camera_dict(cam), (0, 0, 0), 50)
    for fn in (dbg.render_wavefront, dbg.render_megakernel):
        acc = dbg.new_accum(480, cam.img_height)
        fn(acc, 0, 32, 9)
        assert torch.isfinite(acc).all() and 0.1 < float(acc.mean()) / 32 < 1.0
    cs, meta = scenes.heightfield_scene(img_width=320)
    dbg.set_scene(cs)
    dbg.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    acc = dbg.new_accum(meta["width"], meta["height"])
    dbg.render_wavefront(acc, 0, 8, 9)
    assert dbg.debug_error() == 0


def test_a_corrupted_scene_is_caught(dbg):
    """The checks do fire: one leaf reference of the packed tree is redirected to sphere index = sphere count
    (one past the end; still inside the blob, so nothing faults) -> the primitive check records it."""
    cs, meta, _ = golden_fixtures.load_fixture("wavefront_comparison")
    info = dbg.set_scene(cs)
    dbg.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    assert dbg.debug_error() == 0
    nodes = dbg.scene_blob[int(info.off_nodes): int(info.off_nodes) + 64 * int(info.num_inner)].view(torch.int32).view(-1, 16)
    refs = nodes[:, 12:14]
    leaf = (refs < 0).nonzero()[0]
    bad = ~((int(info.num_spheres) << 5) | (2 << 2) | 0)            # sphere #num_spheres, class "Lambert solid"
    refs[leaf[0], leaf[1]] = bad
    torch.cuda.synchronize()
    o, d = dbg.generate_camera_rays(meta["width"], meta["height"], 1, 0)
    dbg.first_hit(o, d, want_pn=False)
    code = dbg.debug_error()
    assert code >> 24 == 2, code                                     # PT_DBG_PRIM
    assert dbg.debug_error() == 0                                    # reading clears it
    dbg.set_scene(cs)                                                # leave a sane scene behind
