"""pt_pack_scene: reference layout -> packed device layout (host side; no GPU needed)."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN_SCENES
from path_tracer_python_b200 import _lib, scenes
import golden_fixtures


def _pack(cs):
    L = _lib.lib()
    hs = _lib.HostScene(cs)
    info = _lib.PtPackedInfo()
    _lib.check(L.pt_pack_scene_info(C.byref(hs.c), C.byref(info)), "info")
    blob = np.zeros(int(info.total_bytes), np.uint8)
    _lib.check(L.pt_pack_scene(C.byref(hs.c), C.byref(info), blob.ctypes.data_as(C.c_void_p)), "pack")
    return info, blob


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_pack_covers_every_primitive_once(name):
    cs, meta, _ = golden_fixtures.load_fixture(name)
    info, blob = _pack(cs)
    n_leaf = int((cs["bvh_prim_idx"] >= 0).sum())
    assert info.num_inner == len(cs["bvh_prim_idx"]) - n_leaf
    nodes = blob[info.off_nodes: info.off_nodes + info.num_inner * 64].view(np.float32).reshape(-1, 16)
    refs = nodes[:, 12:14].copy().view(np.int32).ravel()
    leaves = refs[refs < 0]
    inner = refs[refs >= 0]
    assert sorted(inner.tolist()) == list(range(1, info.num_inner))      # every inner node referenced once, BFS ids
    k = ~leaves
    got = sorted(zip((k & 3).tolist(), (k >> 5).tolist()))
    # bits 2..4: shading class = material bin (6 medium boundary; else 1 emissive, 2 Lambert solid, 3 Lambert textured,
    # 4 metal, 5 dielectric, 7 other; 0 is the kernels' "miss")
    pre = {0: "sphere", 1: "triangle", 2: "quad"}
    for t, i, c in zip((k & 3).tolist(), (k >> 5).tolist(), ((k >> 2) & 7).tolist()):
        med = cs[f"{pre[t]}_is_constant_medium"][i] > 0
        mt, tt = int(cs[f"{pre[t]}_material_type"][i]), int(cs[f"{pre[t]}_texture_type"][i])
        assert c == (6 if med else {0: 2 if tt == 0 else 3, 1: 4, 2: 5, 3: 1}.get(mt, 7))
    want = sorted(zip(cs["bvh_prim_type"][cs["bvh_prim_idx"] >= 0].tolist(), cs["bvh_prim_idx"][cs["bvh_prim_idx"] >= 0].tolist()))
    assert got == want
    # child boxes are the reference's node boxes: root children = nodes 1 and right(0)
    l, r = cs["bvh_left_child"][0], cs["bvh_right_child"][0]
    assert np.array_equal(nodes[0, [0, 4, 8]], cs["bvh_bbox_min"][l]) and np.array_equal(nodes[0, [1, 5, 9]], cs["bvh_bbox_max"][l])
    assert np.array_equal(nodes[0, [2, 6, 10]], cs["bvh_bbox_min"][r]) and np.array_equal(nodes[0, [3, 7, 11]], cs["bvh_bbox_max"][r])
    assert info.max_depth_tree <= 64
    n_prims = info.num_spheres + info.num_quads + info.num_triangles
    ids = blob[info.off_prim_mat: info.off_prim_mat + 4 * n_prims].view(np.uint32)
    assert ids.max() == info.num_materials - 1
    assert info.has_media == int(any(cs[f"{p}_is_constant_medium"].any() for p in ("sphere", "quad", "triangle")))


def test_vol2_materials_deduplicate():
    cs, _, _ = golden_fixtures.load_fixture("vol2_final_scene")
    info, blob = _pack(cs)
    assert info.num_spheres == 1007 and info.num_quads == 2401 and info.num_inner == 3407
    assert info.num_materials == 10          # 3408 primitives share 10 distinct materials
    assert info.num_images == 1 and info.image_w[0] == 1024 and info.image_h[0] == 512
    tex = blob[info.off_images: info.off_images + 1024 * 512 * 4].reshape(512, 1024, 4)
    assert np.array_equal(tex[..., :3], cs["image_0"])


def test_single_primitive_and_empty_scene():
    one = {"sphere_data": np.array([[0, 0, 0, 1]], np.float32), "sphere_material_type": np.zeros(1, np.int32),
           "bvh_bbox_min": np.float32([[-1, -1, -1]]), "bvh_bbox_max": np.float32([[1, 1, 1]]),
           "bvh_left_child": np.int32([-1]), "bvh_right_child": np.int32([-1]), "bvh_prim_type": np.int32([0]),
           "bvh_prim_idx": np.int32([0])}
    for k in ("material_albedo", "material_emit_color", "texture_color1", "texture_color2", "medium_albedo"):
        one["sphere_" + k] = np.zeros((1, 3), np.float32)
    for k in ("material_fuzz", "material_ir", "texture_scale", "medium_density"):
        one["sphere_" + k] = np.zeros(1, np.float32)
    for k in ("texture_type", "texture_image_idx", "is_constant_medium"):
        one["sphere_" + k] = np.zeros(1, np.int32)
    info, blob = _pack(one)
    assert info.num_inner == 1
    refs = blob[info.off_nodes + 48: info.off_nodes + 56].view(np.int32)
    # leaf ref for sphere 0 (class 2 = Lambertian, solid colour) in BOTH slots: every packed node has two real children
    assert refs[0] == ~(2 << 2) and refs[1] == refs[0]
    info, blob = _pack({})
    assert info.num_inner == 0 and info.total_bytes >= 256


def test_malformed_scenes_are_rejected_not_dereferenced():
    """`compiled_scene=` is public API.  A missing key or a short per-primitive array
    raises in HostScene; index ranges the Python side cannot see are rejected by the C packer with
    PT_ERR_INVALID_ARG (never a NULL / out-of-bounds dereference)."""
    cs, _, _ = golden_fixtures.load_fixture("wavefront_comparison")
    for drop in ("sphere_material_type", "sphere_is_constant_medium", "bvh_prim_type", "bvh_left_child"):
        bad = {k: v for k, v in cs.items() if k != drop}
        with pytest.raises(_lib.B200PTError, match=drop):
            _lib.HostScene(bad)
    bad = dict(cs, sphere_material_fuzz=cs["sphere_material_fuzz"][:-1])
    with pytest.raises(_lib.B200PTError, match="rows"):
        _lib.HostScene(bad)
    bad = dict(cs, sphere_material_albedo=cs["sphere_material_albedo"][:, :2])
    with pytest.raises(_lib.B200PTError, match="shape"):
        _lib.HostScene(bad)
    L = _lib.lib()
    leaf = int(np.where(cs["bvh_prim_idx"] >= 0)[0][0])
    for key, value, msg in (("bvh_prim_type", 3, "bvh_prim_type"), ("bvh_prim_idx", 10 ** 6, "bvh_prim_idx"),
                            ("bvh_prim_type", 2, "bvh_prim_idx")):          # type 2 = quad: the scene has none
        a = cs[key].copy()
        a[leaf] = value
        hs = _lib.HostScene(dict(cs, **{key: a}))
        info = _lib.PtPackedInfo()
        assert L.pt_pack_scene_info(C.byref(hs.c), C.byref(info)) == -1      # PT_ERR_INVALID_ARG
        assert msg in L.pt_last_error().decode()
    mt = cs["sphere_material_type"].copy()
    mt[0] = 9
    hs = _lib.HostScene(dict(cs, sphere_material_type=mt))
    assert L.pt_pack_scene_info(C.byref(hs.c), C.byref(_lib.PtPackedInfo())) == -1
    # an image-textured Lambertian whose image index points past the image table
    z, _, _ = golden_fixtures.load_fixture("feature_zoo")
    idx = z["sphere_texture_image_idx"].copy()
    idx[z["sphere_texture_type"] == 2] = 5
    hs = _lib.HostScene(dict(z, sphere_texture_image_idx=idx))
    assert L.pt_pack_scene_info(C.byref(hs.c), C.byref(_lib.PtPackedInfo())) == -1
