"""BVH compile step: replaces ``compile_bvh`` (src/render_server/taichi_renderer/bvh_compiler.py:132)
and the pure-Python ``SAHBVHBuilder`` (sah_bvh_builder.py:81-485, 2.4 s for the 3408 primitives of
vol2_final_scene) with the native builder in libb200pt (pt_bvh_build_sah), which reproduces the
reference's tree bit for bit (tests/test_bvh_builder.py) in milliseconds.

Primitive boxes / centroids are prepared here with vectorised float32 numpy, following
add_sphere / add_quad / add_triangle (sah_bvh_builder.py:104-165) operation by operation.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

PRIM_SPHERE, PRIM_TRIANGLE, PRIM_QUAD = 0, 1, 2


def _pad_to_minimums(mn: np.ndarray, mx: np.ndarray):
    """AABB.pad_to_minimums (sah_bvh_builder.py:31-39), float32."""
    delta = np.float32(0.0001)
    thin = (mx - mn) < delta
    mid = (mn + mx) / np.float32(2.0)
    half = np.float32(0.0001 / 2.0)
    mn = np.where(thin, mid - half, mn)
    mx = np.where(thin, mid + half, mx)
    return mn.astype(np.float32), mx.astype(np.float32)


def primitive_bounds(cs: dict):
    """-> (bbox_min, bbox_max, centroid, prim_type, prim_idx) in builder insertion order:
    spheres, quads, triangles (sah_bvh_builder.py:461-479)."""
    parts = []
    sd = np.asarray(cs.get("sphere_data", np.zeros((0, 4))), np.float32)
    if len(sd):
        c, r = sd[:, :3], sd[:, 3:4]
        lo, hi = c - r, c + r
        if int(cs.get("full_features", 0)) and "sphere_motion" in cs:
            # full feature mode: box = union of the boxes at both ends of the motion (core/sphere.py:27-32)
            c1 = c + np.asarray(cs["sphere_motion"], np.float32)
            lo, hi = np.minimum(lo, c1 - r), np.maximum(hi, c1 + r)
        parts.append((lo, hi, c.copy(), PRIM_SPHERE, len(sd)))
    Q = np.asarray(cs.get("quad_Q", np.zeros((0, 3))), np.float32)
    if len(Q):
        u, v = np.asarray(cs["quad_u"], np.float32), np.asarray(cs["quad_v"], np.float32)
        corners = np.stack([Q, Q + u, Q + v, Q + u + v])
        mn, mx = _pad_to_minimums(corners.min(axis=0), corners.max(axis=0))
        cen = Q + np.float32(0.5) * u + np.float32(0.5) * v
        parts.append((mn, mx, cen, PRIM_QUAD, len(Q)))
    v0 = np.asarray(cs.get("triangle_v0", np.zeros((0, 3))), np.float32)
    if len(v0):
        v1, v2 = np.asarray(cs["triangle_v1"], np.float32), np.asarray(cs["triangle_v2"], np.float32)
        verts = np.stack([v0, v1, v2])
        mn, mx = _pad_to_minimums(verts.min(axis=0), verts.max(axis=0))
        cen = (v0 + v1 + v2) / np.float32(3.0)
        parts.append((mn, mx, cen, PRIM_TRIANGLE, len(v0)))
    if not parts:
        z = np.zeros((0, 3), np.float32)
        return z, z, z, np.zeros(0, np.int32), np.zeros(0, np.int32)
    bmin = np.ascontiguousarray(np.concatenate([p[0] for p in parts]), np.float32)
    bmax = np.ascontiguousarray(np.concatenate([p[1] for p in parts]), np.float32)
    cen = np.ascontiguousarray(np.concatenate([p[2] for p in parts]), np.float32)
    ptype = np.concatenate([np.full(p[4], p[3], np.int32) for p in parts])
    pidx = np.concatenate([np.arange(p[4], dtype=np.int32) for p in parts])
    return bmin, bmax, cen, ptype, pidx


def build_sah(bmin, bmax, cen, ptype, pidx) -> dict:
    n = len(ptype)
    cap = max(2 * n - 1, 1)
    out = {
        "bvh_bbox_min": np.zeros((cap, 3), np.float32), "bvh_bbox_max": np.zeros((cap, 3), np.float32),
        "bvh_left_child": np.full(cap, -1, np.int32), "bvh_right_child": np.full(cap, -1, np.int32),
        "bvh_parent": np.full(cap, -1, np.int32), "bvh_prim_type": np.zeros(cap, np.int32),
        "bvh_prim_idx": np.full(cap, -1, np.int32),
    }
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = _lib.lib().pt_bvh_build_sah(
        C.c_int32(n), p(bmin), p(bmax), p(cen), p(np.ascontiguousarray(ptype, np.int32)),
        p(np.ascontiguousarray(pidx, np.int32)), p(out["bvh_bbox_min"]), p(out["bvh_bbox_max"]),
        p(out["bvh_left_child"]), p(out["bvh_right_child"]), p(out["bvh_parent"]), p(out["bvh_prim_type"]),
        p(out["bvh_prim_idx"]))
    if rc < 0:
        _lib.check(rc, "pt_bvh_build_sah")
    return {k: v[:rc] for k, v in out.items()}


def compile_bvh(cs: dict) -> dict:
    """CompiledScene geometry -> the 7 node arrays of compile_bvh (bvh_compiler.py:132-168)."""
    return build_sah(*primitive_bounds(cs))
