"""Native binned-SAH builder (pt_bvh_build_sah) == the reference's SAHBVHBuilder, bit for bit."""
import numpy as np
import pytest

from conftest import GOLDEN_SCENES
from path_tracer_python_b200 import bvh, scenes
import golden_fixtures

KEYS = ["bvh_bbox_min", "bvh_bbox_max", "bvh_left_child", "bvh_right_child", "bvh_parent", "bvh_prim_type",
        "bvh_prim_idx"]


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_matches_reference_tree(name):
    cs, meta, _ = golden_fixtures.load_fixture(name)
    mine = bvh.compile_bvh(cs)
    for k in KEYS:
        assert mine[k].dtype == cs[k].dtype
        assert np.array_equal(mine[k], cs[k]), k


def test_empty_and_single():
    z3 = np.zeros((0, 3), np.float32)
    out = bvh.build_sah(z3, z3, z3, np.zeros(0, np.int32), np.zeros(0, np.int32))
    assert len(out["bvh_prim_idx"]) == 0
    one = bvh.compile_bvh({"sphere_data": np.array([[1, 2, 3, 0.5]], np.float32)})
    assert len(one["bvh_prim_idx"]) == 1 and one["bvh_prim_idx"][0] == 0 and one["bvh_left_child"][0] == -1
    assert np.array_equal(one["bvh_bbox_min"][0], np.float32([0.5, 1.5, 2.5]))


def test_coincident_centroids_fall_back_to_median():
    # all centroids equal -> no SAH split exists (sah_bvh_builder.py:263, :330-334) -> median split
    n = 9
    sd = np.tile(np.array([[0, 0, 0, 1]], np.float32), (n, 1))
    sd[:, 3] = np.arange(1, n + 1)
    out = bvh.compile_bvh({"sphere_data": sd})
    assert len(out["bvh_prim_idx"]) == 2 * n - 1
    leaves = out["bvh_prim_idx"][out["bvh_prim_idx"] >= 0]
    assert sorted(leaves.tolist()) == list(range(n))


def test_structure_invariants_random():
    rng = np.random.default_rng(5)
    n = 5000
    sd = np.concatenate([rng.uniform(-100, 100, (n, 3)), rng.uniform(0.1, 3, (n, 1))], axis=1).astype(np.float32)
    out = bvh.compile_bvh({"sphere_data": sd})
    N = len(out["bvh_prim_idx"])
    assert N == 2 * n - 1
    leaf = out["bvh_prim_idx"] >= 0
    assert sorted(out["bvh_prim_idx"][leaf].tolist()) == list(range(n))
    inner = np.where(~leaf)[0]
    assert np.all(out["bvh_left_child"][inner] == inner + 1)          # DFS pre-order: left child follows
    l, r = out["bvh_left_child"][inner], out["bvh_right_child"][inner]
    assert np.array_equal(np.minimum(out["bvh_bbox_min"][l], out["bvh_bbox_min"][r]), out["bvh_bbox_min"][inner])
    assert np.array_equal(np.maximum(out["bvh_bbox_max"][l], out["bvh_bbox_max"][r]), out["bvh_bbox_max"][inner])
    assert np.all(out["bvh_parent"][l] == inner) and np.all(out["bvh_parent"][r] == inner)
