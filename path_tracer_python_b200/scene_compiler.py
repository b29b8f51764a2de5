"""Host compile step: scene object graph -> flat float32/int32 arrays ("CompiledScene").

Mirrors the *output contract* of the reference's ``compile_scene``
(src/render_server/taichi_renderer/scene_compiler.py:931-965): the same arrays, dtypes, ordering
and the same quirks (object de-duplication by identity :38-50, sphere centre baked at time 0
:161-162, medium boundaries flagged through a registry :854-928, unknown materials become
0.8-grey Lambertian :406-417, image textures found only on primitives that are not inside a
medium or a mesh :829-845).  The arrays are returned as one flat dict -- the FFI payload of
libb200pt (include/b200pt.h PtHostScene) -- instead of the reference's 11-tuple.

The scene classes are duck-typed by class name, so both the reference's ``src/core`` objects and
the light-weight stand-ins of ``path_tracer_python_b200.scene`` are accepted.
tests/test_scene_compile.py checks array-for-array equality with the reference's own compiler.
"""
from __future__ import annotations

import numpy as np

PRIM_SPHERE, PRIM_TRIANGLE, PRIM_QUAD = 0, 1, 2
MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, MAT_EMISSIVE, MAT_ISOTROPIC = 0, 1, 2, 3, 4
MAT_SSS_SIMPLE, MAT_SSS_VOLUMETRIC = 5, 6        # `full` feature mode only
TEX_SOLID, TEX_CHECKER, TEX_IMAGE, TEX_NOISE = 0, 1, 2, 3

_PREFIX = {PRIM_SPHERE: "sphere", PRIM_TRIANGLE: "triangle", PRIM_QUAD: "quad"}


def _kind(obj) -> str:
    return type(obj).__name__


class _Origin:
    x = y = z = 0.0


def _xyz(v):
    return [v.x, v.y, v.z]


def _walk(obj, visit, through_media: bool, through_mesh: bool):
    """Depth-first walk in the reference's order: lists front to back, bvh nodes left then right."""
    k = _kind(obj)
    if k == "constant_medium":
        if through_media:
            _walk(obj.boundary, visit, through_media, through_mesh)
    elif k == "mesh":
        if through_mesh:
            if hasattr(obj, "tri_arrays"):
                visit(obj)                      # array-backed mesh: one chunk, no per-triangle objects
            else:
                for tri in obj.triangles:
                    visit(tri)
    elif k in ("Sphere", "quad", "triangle"):
        visit(obj)
    elif k == "hittable_list":
        for item in obj.objects:
            _walk(item, visit, through_media, through_mesh)
    elif k == "bvh_node":
        if obj.left is not None:
            _walk(obj.left, visit, through_media, through_mesh)
        if obj.right is not None:
            _walk(obj.right, visit, through_media, through_mesh)


def extract_primitives(world):
    """-> (spheres, quads, triangles), each unique by identity, in traversal order."""
    out = {"Sphere": [], "quad": [], "triangle": []}
    seen = set()

    def visit(p):
        if id(p) not in seen:
            seen.add(id(p))
            out["triangle" if _kind(p) == "mesh" else _kind(p)].append(p)

    _walk(world, visit, through_media=True, through_mesh=True)
    return out["Sphere"], out["quad"], out["triangle"]


def _material_of(prim):
    return prim.material if hasattr(prim, "material") else prim.mat


def build_image_registry(world):
    """id(texture) -> index, in order of first appearance (scene_compiler.py:812-851)."""
    textures = {}

    def visit(p):
        mat = _material_of(p)
        tex = getattr(mat, "tex", None)
        if tex is not None and _kind(tex) == "image_texture":
            textures[id(tex)] = tex

    _walk(world, visit, through_media=False, through_mesh=False)
    lst = list(textures.values())
    return {id(t): i for i, t in enumerate(lst)}, lst


def build_medium_registry(world):
    """id(boundary primitive) -> (density, albedo) (scene_compiler.py:854-928)."""
    reg = {}

    def find(obj):
        k = _kind(obj)
        if k == "constant_medium":
            density = -1.0 / obj.neg_inv_density if obj.neg_inv_density != 0 else 0.01
            albedo = (1.0, 1.0, 1.0)
            tex = getattr(obj.phase_function, "tex", None)
            if tex is not None:
                try:
                    c = tex.value(0, 0, _Origin())
                    albedo = (c.x, c.y, c.z)
                except Exception:
                    pass
            _walk(obj.boundary, lambda p: reg.__setitem__(id(p), (density, albedo)), False, False)
        elif k == "hittable_list":
            for item in obj.objects:
                find(item)
        elif k == "bvh_node":
            if obj.left is not None:
                find(obj.left)
            if obj.right is not None:
                find(obj.right)

    find(world)
    return reg


def _material_arrays(prims, ref_point_of, image_registry, medium_registry, full=False):
    n = len(prims)
    a = {
        "material_type": np.zeros(n, np.int32), "material_albedo": np.zeros((n, 3), np.float32),
        "material_fuzz": np.zeros(n, np.float32), "material_ir": np.zeros(n, np.float32),
        "material_emit_color": np.zeros((n, 3), np.float32), "texture_type": np.zeros(n, np.int32),
        "texture_scale": np.ones(n, np.float32), "texture_color1": np.zeros((n, 3), np.float32),
        "texture_color2": np.zeros((n, 3), np.float32), "texture_image_idx": np.full(n, -1, np.int32),
        "is_constant_medium": np.zeros(n, np.int32), "medium_density": np.zeros(n, np.float32),
        "medium_albedo": np.zeros((n, 3), np.float32),
    }
    for i, prim in enumerate(prims):
        mat = _material_of(prim)
        ref = ref_point_of(prim)
        mk = _kind(mat)
        fuzz, ir = 0.0, 1.0
        tex_type, scale, img = TEX_SOLID, 1.0, -1
        albedo = c1 = c2 = (0.8, 0.8, 0.8)
        emit = (0.0, 0.0, 0.0)
        mtype = MAT_LAMBERTIAN
        if mk == "lambertian":
            tk = _kind(mat.tex)
            albedo = (1.0, 1.0, 1.0)
            if tk == "checker_texture":
                tex_type, scale = TEX_CHECKER, 1.0 / mat.tex.inv_scale
                c1, c2 = _xyz(mat.tex.even.value(0, 0, ref)), _xyz(mat.tex.odd.value(0, 0, ref))
            elif tk == "image_texture":
                tex_type, img = TEX_IMAGE, image_registry.get(id(mat.tex), -1)
                c1 = c2 = (1.0, 1.0, 1.0)
            elif tk == "noise_texture":
                tex_type, scale = TEX_NOISE, mat.tex.scale
                c1 = c2 = (0.5, 0.5, 0.5)
            else:
                try:
                    albedo = c1 = c2 = _xyz(mat.tex.value(0, 0, ref))
                except Exception:
                    albedo = c1 = c2 = (0.8, 0.8, 0.8)
        elif mk == "metal":
            mtype, fuzz = MAT_METAL, mat.fuzz
            albedo = c1 = c2 = _xyz(mat.albedo)
        elif mk == "dielectric":
            mtype, ir = MAT_DIELECTRIC, mat.ir
            albedo = c1 = c2 = (1.0, 1.0, 1.0)
        elif mk == "diffuse_light":
            mtype = MAT_EMISSIVE
            try:
                emit = _xyz(mat.tex.value(0, 0, ref))
            except Exception:
                emit = (1.0, 1.0, 1.0)
            albedo = c1 = c2 = (0.0, 0.0, 0.0)
        elif full and mk == "subsurface_simple":
            # full feature mode: parameters ride in the existing arrays (fuzz = scatter distance)
            mtype, fuzz = MAT_SSS_SIMPLE, mat.scatter_distance
            albedo = c1 = c2 = _xyz(mat.albedo)
        elif full and mk == "subsurface_volumetric":
            # fuzz = sigma_s, ir = sigma_a, texture_scale = g
            mtype, fuzz, ir, scale = MAT_SSS_VOLUMETRIC, mat.sigma_s, mat.sigma_a, mat.g
            albedo = c1 = c2 = _xyz(mat.albedo)
        # anything else (isotropic, subsurface_* in compat mode, ...) stays 0.8-grey Lambertian (:406-417)
        a["material_type"][i] = mtype
        a["material_albedo"][i] = albedo
        a["material_fuzz"][i] = fuzz
        a["material_ir"][i] = ir
        a["material_emit_color"][i] = emit
        a["texture_type"][i] = tex_type
        a["texture_scale"][i] = scale
        a["texture_color1"][i] = c1
        a["texture_color2"][i] = c2
        a["texture_image_idx"][i] = img
        med = medium_registry.get(id(prim))
        if med is not None:
            a["is_constant_medium"][i] = 1
            a["medium_density"][i] = med[0]
            a["medium_albedo"][i] = med[1]
    return a


class _MeshTri:
    """Stand-in for "a triangle of this mesh" when evaluating the mesh's material once."""

    def __init__(self, m):
        self.mat = m.mat
        a = m.tri_arrays()[0][0]
        self.v0 = type("P", (), {"x": float(a[0]), "y": float(a[1]), "z": float(a[2])})()


def compile_scene(world, features: str = "taichi_compat") -> dict:
    """world -> CompiledScene dict (geometry, materials, image textures; no BVH, no Perlin).

    features="taichi_compat" (default) reproduces the reference's Taichi path exactly: sphere
    motion baked at time 0, subsurface materials -> grey Lambertian.  "nee" / "full+nee" add next-event
    estimation at Lambertian hits (same expected image, less noise; the technique the reference's docs list
    as missing, docs/__optimizations.md:7-10).  features="full" additionally
    exports what only the reference's pure-Python tracer renders: `sphere_motion` (centre(1) -
    centre(0), core/sphere.py:22-36) and the two subsurface materials (core/material.py:144-276)."""
    if features not in ("taichi_compat", "full", "nee", "full+nee"):
        raise ValueError("features must be 'taichi_compat', 'full', 'nee' or 'full+nee'")
    full = features in ("full", "full+nee")
    nee = features in ("nee", "full+nee")
    image_registry, image_list = build_image_registry(world)
    medium_registry = build_medium_registry(world)
    spheres, quads, triangles = extract_primitives(world)
    cs = {}
    sd = np.zeros((len(spheres), 4), np.float32)
    for i, s in enumerate(spheres):
        c = s.center.at(0.0)                      # motion baked at time 0 (:161-162)
        sd[i] = (c.x, c.y, c.z, s.radius)
    cs["sphere_data"] = sd

    def vec_array(objs, attr):
        out = np.zeros((len(objs), 3), np.float32)
        for i, o in enumerate(objs):
            out[i] = _xyz(getattr(o, attr))
        return out

    for attr in ("Q", "u", "v", "normal", "w"):
        cs[f"quad_{attr}"] = vec_array(quads, attr)
    cs["quad_D"] = np.array([q.D for q in quads], np.float32).reshape(len(quads))
    # triangles: single objects and array-backed mesh chunks, in traversal order; edges and the unit
    # normal are computed in float64 like core/triangle.py:23-30, then narrowed to float32
    chunks, tri_ref, tri_mat = [], [], []
    for t in triangles:
        if _kind(t) == "mesh":
            v0, v1, v2 = t.tri_arrays()
            chunks.append((np.asarray(v0, np.float64), np.asarray(v1, np.float64), np.asarray(v2, np.float64)))
            tri_mat.append((t.mat, len(v0)))
        else:
            chunks.append(tuple(np.array([_xyz(getattr(t, a))], np.float64) for a in ("v0", "v1", "v2")))
            tri_mat.append((t, 1))
    if chunks:
        v0, v1, v2 = (np.concatenate([c[k] for c in chunks]) for k in range(3))
    else:
        v0 = v1 = v2 = np.zeros((0, 3), np.float64)
    e1, e2 = v1 - v0, v2 - v0
    nrm = np.cross(e1, e2)
    with np.errstate(invalid="ignore", divide="ignore"):
        nrm = nrm / np.sqrt((nrm ** 2).sum(axis=1, keepdims=True))
    # single triangle objects carry their own (identically computed) attributes: use them verbatim
    row = 0
    for (obj, n) in tri_mat:
        if n == 1 and _kind(obj) == "triangle":
            e1[row], e2[row], nrm[row] = _xyz(obj.edge1), _xyz(obj.edge2), _xyz(obj.normal)
        row += n
    for name, arr in (("v0", v0), ("v1", v1), ("v2", v2), ("edge1", e1), ("edge2", e2), ("normal", nrm)):
        cs[f"triangle_{name}"] = np.ascontiguousarray(arr, np.float32)
    per_type = ((PRIM_SPHERE, spheres, lambda s: s.center.at(0.0)), (PRIM_QUAD, quads, lambda q: q.Q),
                (PRIM_TRIANGLE, triangles, lambda t: t.v0))
    for pt, prims, ref_of in per_type:
        if pt == PRIM_TRIANGLE and any(_kind(t) == "mesh" for t in prims):
            # one material row per mesh, repeated for its triangles
            parts = []
            for t in prims:
                if _kind(t) == "mesh":
                    proxy = _MeshTri(t)
                    rows = _material_arrays([proxy], lambda q: q.v0, image_registry, medium_registry, full)
                    parts.append({k: np.repeat(v, len(t.tri_arrays()[0]), axis=0) for k, v in rows.items()})
                else:
                    parts.append(_material_arrays([t], ref_of, image_registry, medium_registry, full))
            for k in parts[0]:
                cs[f"triangle_{k}"] = np.concatenate([p[k] for p in parts])
            continue
        for k, v in _material_arrays(prims, ref_of, image_registry, medium_registry, full).items():
            cs[f"{_PREFIX[pt]}_{k}"] = v
    if full:
        mo = np.zeros((len(spheres), 3), np.float32)
        for i, sp in enumerate(spheres):
            c0, c1 = sp.center.at(0.0), sp.center.at(1.0)
            mo[i] = (c1.x - c0.x, c1.y - c0.y, c1.z - c0.z)
        cs["sphere_motion"] = mo
        cs["full_features"] = 1
    if nee:
        # next-event estimation at Lambertian hits (bit 1 of the feature word); the kernels' feature-mode
        # instantiations always read the motion table, so a compat-material scene gets one of zeros
        cs.setdefault("sphere_motion", np.zeros((len(spheres), 3), np.float32))
        cs["full_features"] = int(cs.get("full_features", 0)) | 2
    cs["num_images"] = len(image_list)
    for i, tex in enumerate(image_list):
        cs[f"image_{i}"] = image_to_u8(tex)
    return cs


def image_to_u8(tex) -> np.ndarray:
    """8-bit RGB texels of an image_texture (util/rtw_image.py:66 keeps float32 = u8/255)."""
    img = tex.image
    if getattr(img, "bdata8", None) is not None:
        return np.ascontiguousarray(img.bdata8, np.uint8)
    f = np.asarray(img.fdata, np.float32)
    u8 = np.rint(f * 255.0).astype(np.uint8)
    if not np.array_equal(u8.astype(np.float32) / np.float32(255.0), f):
        raise ValueError("image texture is not an 8-bit image divided by 255; cannot be stored as RGBA8")
    return u8


def perlin_tables(rng=None) -> dict:
    """Gradient vectors + permutations drawn the way core/perlin.py:7-55 draws them from the global
    `random` stream (256 x 3 uniform(-1,1), then three Fisher-Yates shuffles using randint(0, i)),
    so a seeded session produces the very tables the reference renderer would upload."""
    import random as _random

    r = rng or _random
    vec = np.array([[r.uniform(-1, 1), r.uniform(-1, 1), r.uniform(-1, 1)] for _ in range(256)], np.float32)
    perms = []
    for _ in range(3):
        p = list(range(256))
        for i in range(255, 0, -1):
            t = r.randint(0, i)
            p[i], p[t] = p[t], p[i]
        perms.append(np.array(p, np.int32))
    return {"perlin_randvec": vec, "perlin_perm_x": perms[0], "perlin_perm_y": perms[1], "perlin_perm_z": perms[2]}
