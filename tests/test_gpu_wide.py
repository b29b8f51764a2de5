"""The 4-wide BVH build (libb200pt_wide.so, -DPT_WIDE=1): the packer collapses the reference's binary SAH tree into
128-byte nodes with up to four children (child boxes unchanged), the kernels test four boxes per visit, sort the hit
children with a 5-comparator network and push up to three.  A ray reaches exactly the leaves whose own boxes it
passes, as in the binary tree, so every result is bit-identical -- checked here on the golden suite; it is not the
default because it measured 5 % slower on config 2."""
import ctypes as C

import numpy as np
import pytest

import golden_fixtures
from conftest import GOLDEN_SCENES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wide():
    from path_tracer_python_b200.engine import Engine
    if not Engine.has_numerics("wide"):
        pytest.fail("libb200pt_wide.so is not built (python -m path_tracer_python_b200.build)")
    e = Engine(0, "wide")
    assert b"4-wide" in e.L.pt_version()
    return e


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_wide_tree_gives_the_reference_bits(wide, name):
    cs, meta, out = golden_fixtures.load_fixture(name)
    info = wide.set_scene(cs)
    assert info.node_bytes == 128
    n_leaf = int((cs["bvh_prim_idx"] >= 0).sum())
    assert (n_leaf - 1 + 2) // 3 <= info.num_inner <= max(1, n_leaf - 1)        # between a full 4-ary and the binary tree
    wide.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    W, H, spp, seed = meta["width"], meta["height"], meta["spp"], meta["rng_seed"]
    acc = wide.new_accum(W, H)
    wide.render_megakernel(acc, 0, spp, seed)
    assert np.array_equal(acc.cpu().numpy(), out["accum_megakernel"])
    acc = wide.new_accum(W, H)
    wide.render_wavefront(acc, 0, spp, seed)
    assert np.array_equal(acc.cpu().numpy(), out["accum_wavefront"])
    o, d = wide.generate_camera_rays(W, H, seed, 0)
    t, prim, p, n = (x.cpu().numpy() for x in wide.first_hit(o, d))
    valid = out["hit_valid"] == 1
    assert np.array_equal(prim >= 0, valid) and np.array_equal(t[valid], out["hit_t"][valid])
    # every leaf of the reference tree appears exactly once among the wide nodes' child references
    nodes = wide.scene_blob[int(info.off_nodes): int(info.off_nodes) + 128 * int(info.num_inner)].cpu().numpy().view(np.int32).reshape(-1, 32)
    refs = nodes[:, 24:28].ravel()
    leaves = ~refs[(refs < 0) & (refs != -2 ** 31)]
    got = sorted(zip((leaves & 3).tolist(), (leaves >> 5).tolist()))
    want = sorted(zip(cs["bvh_prim_type"][cs["bvh_prim_idx"] >= 0].tolist(), cs["bvh_prim_idx"][cs["bvh_prim_idx"] >= 0].tolist()))
    assert got == want


def test_the_two_layouts_do_not_mix(wide):
    """A blob packed for the binary layout is refused by the wide library (and vice versa): node_bytes is checked."""
    from path_tracer_python_b200 import _lib
    cs, _, _ = golden_fixtures.load_fixture("wavefront_comparison")
    hs = _lib.HostScene(cs)
    info = _lib.PtPackedInfo()
    _lib.check(_lib.lib("strict").pt_pack_scene_info(C.byref(hs.c), C.byref(info)), "info")
    assert info.node_bytes == 64
    assert wide.L.pt_set_scene(wide.ctx, C.c_void_p(wide.scene_blob.data_ptr()), C.byref(info)) == -1
    assert b"node layout" in wide.L.pt_last_error()
