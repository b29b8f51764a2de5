"""ctypes binding of libb200pt.so (include/b200pt.h).  No CPU fallback: if the CUDA library is
missing or fails to load, importing the renderer raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PT_LIB_PATH") or os.path.join(_HERE, "libb200pt.so")   # PT_LIB_PATH: A/B builds
# numerics="fast": the same sources built with -DPT_FAST=1 --fmad=true (build.py); never the parity build
# numerics="debug": strict numerics with every index checked (-DPT_DEBUG_BOUNDS=1), for the parity suite only
LIB_PATHS = {"strict": LIB_PATH, "fast": os.environ.get("PT_FAST_LIB_PATH") or os.path.join(_HERE, "libb200pt_fast.so"),
             "debug": os.path.join(_HERE, "libb200pt_dbg.so"),
             "wide": os.path.join(_HERE, "libb200pt_wide.so")}     # strict numerics over the 4-wide collapsed BVH (-DPT_WIDE=1)

PT_MAX_IMAGES = 16
PRIM_SPHERE, PRIM_TRIANGLE, PRIM_QUAD = 0, 1, 2
PRIM_PREFIX = {PRIM_SPHERE: "sphere", PRIM_TRIANGLE: "triangle", PRIM_QUAD: "quad"}


class PtHostMaterials(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "material_type", "albedo", "fuzz", "ir", "emit", "texture_type", "texture_scale", "color1", "color2",
        "image_idx", "is_medium", "medium_density", "medium_albedo")]


class PtHostScene(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in ("num_spheres", "num_quads", "num_triangles", "num_nodes", "num_images")]
        + [(n, C.c_void_p) for n in ("sphere_data", "quad_Q", "quad_u", "quad_v", "quad_normal", "quad_D", "quad_w",
                                     "tri_v0", "tri_edge1", "tri_edge2", "tri_normal")]
        + [("mat", PtHostMaterials * 3)]
        + [(n, C.c_void_p) for n in ("bvh_bbox_min", "bvh_bbox_max", "bvh_left", "bvh_right", "bvh_prim_type",
                                     "bvh_prim_idx", "perlin_randvec", "perm_x", "perm_y", "perm_z")]
        + [("images", C.c_void_p * PT_MAX_IMAGES), ("image_w", C.c_int32 * PT_MAX_IMAGES),
           ("image_h", C.c_int32 * PT_MAX_IMAGES)]
        + [("sphere_motion", C.c_void_p), ("full_features", C.c_int32)]
    )


class PtPackedInfo(C.Structure):
    _fields_ = (
        [(n, C.c_uint64) for n in ("total_bytes", "off_nodes", "off_spheres", "off_quads", "off_tris", "off_prim_mat",
                                   "off_materials", "off_perlin", "off_images")]
        + [("image_off", C.c_uint64 * PT_MAX_IMAGES), ("image_w", C.c_int32 * PT_MAX_IMAGES),
           ("image_h", C.c_int32 * PT_MAX_IMAGES)]
        + [(n, C.c_int32) for n in ("num_inner", "num_spheres", "num_quads", "num_triangles", "num_materials",
                                    "num_images", "max_depth_tree", "has_media")]
        + [("off_sphere_motion", C.c_uint64), ("full_features", C.c_int32), ("node_bytes", C.c_int32),
           ("off_lights", C.c_uint64), ("num_lights", C.c_int32), ("nee_r2", C.c_float)]
    )


class PtCamera(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("pixel00", C.c_float * 3), ("delta_u", C.c_float * 3),
                ("delta_v", C.c_float * 3), ("defocus_disk_u", C.c_float * 3), ("defocus_disk_v", C.c_float * 3),
                ("defocus_angle", C.c_float), ("background", C.c_float * 3), ("max_depth", C.c_int32),
                ("jitter", C.c_int32)]


class PtStageTime(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("wave", C.c_int32), ("ms", C.c_float)]


STAGE_NAMES = ("reset", "generate", "primary", "extend", "bin", "shade", "tail", "resolve")


class PtCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "box_tests", "sphere_tests", "triangle_tests", "quad_tests",
                                          "paths", "bounces")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


# every symbol include/b200pt.h declares (tests/test_abi.py checks the library exports them all)
EXPORTS = [
    "pt_last_error", "pt_version", "pt_pack_scene_info", "pt_pack_scene", "pt_bvh_build_sah", "pt_create",
    "pt_destroy", "pt_set_option", "pt_set_row_interleave", "pt_set_scene", "pt_set_camera", "pt_render_megakernel",
    "pt_wavefront_workspace_bytes", "pt_wavefront_workspace_bytes_for", "pt_render_wavefront", "pt_wf_begin", "pt_wf_generate", "pt_wf_extend", "pt_wf_shade", "pt_wf_compact",
    "pt_wf_read_rays", "pt_wf_read_hits", "pt_wf_resolve", "pt_generate_camera_rays", "pt_first_hit",
    "pt_math_eval", "pt_probe_gather", "pt_tonemap", "pt_get_debug_error", "pt_enable_counters", "pt_get_counters", "pt_enable_stage_timing", "pt_get_stage_times", "pt_get_stage_counters",
]

_libs = {}


class B200PTError(RuntimeError):
    pass


def has_numerics(numerics: str) -> bool:
    return numerics in LIB_PATHS and os.path.exists(LIB_PATHS[numerics])


def lib(numerics: str = "strict") -> C.CDLL:
    """Load libb200pt.so (or the fast-numerics variant).  Raises if it has not been built
    (python -m path_tracer_python_b200.build)."""
    if numerics in _libs:
        return _libs[numerics]
    if numerics not in LIB_PATHS:
        raise B200PTError(f"unknown numerics {numerics!r} (strict | fast | debug | wide)")
    path = LIB_PATHS[numerics]
    if not os.path.exists(path):
        raise B200PTError(
            f"{path} not found: the CUDA extension is not built. Run `python -m path_tracer_python_b200.build` "
            "(there is no CPU fallback).")
    L = C.CDLL(path)
    L.pt_last_error.restype = C.c_char_p
    L.pt_version.restype = C.c_char_p
    L.pt_wavefront_workspace_bytes.restype = C.c_size_t
    L.pt_wavefront_workspace_bytes.argtypes = [C.c_int, C.c_int]
    L.pt_wavefront_workspace_bytes_for.restype = C.c_size_t
    L.pt_wavefront_workspace_bytes_for.argtypes = [C.c_int, C.c_int, C.c_int]
    L.pt_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.pt_destroy.argtypes = [C.c_void_p]
    L.pt_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.pt_set_row_interleave.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.pt_set_scene.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(PtPackedInfo)]
    L.pt_set_camera.argtypes = [C.c_void_p, C.POINTER(PtCamera)]
    L.pt_pack_scene_info.argtypes = [C.POINTER(PtHostScene), C.POINTER(PtPackedInfo)]
    L.pt_pack_scene.argtypes = [C.POINTER(PtHostScene), C.POINTER(PtPackedInfo), C.c_void_p]
    L.pt_render_megakernel.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_uint32, C.c_void_p]
    L.pt_render_wavefront.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p]
    L.pt_wf_begin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_size_t]
    L.pt_wf_generate.argtypes = [C.c_void_p, C.c_void_p]
    L.pt_wf_extend.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.pt_wf_shade.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.pt_wf_compact.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]
    L.pt_wf_read_rays.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
    L.pt_wf_read_hits.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    L.pt_wf_resolve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.pt_generate_camera_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
    L.pt_first_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pt_math_eval.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pt_probe_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                  C.c_void_p]
    L.pt_tonemap.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.pt_get_debug_error.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    L.pt_enable_counters.argtypes = [C.c_void_p, C.c_int]
    L.pt_get_counters.argtypes = [C.c_void_p, C.POINTER(PtCounters), C.c_int]
    L.pt_enable_stage_timing.argtypes = [C.c_void_p, C.c_int]
    L.pt_get_stage_times.argtypes = [C.c_void_p, C.POINTER(PtStageTime), C.c_int, C.POINTER(C.c_int)]
    L.pt_get_stage_counters.argtypes = [C.c_void_p, C.POINTER(PtCounters), C.c_int, C.POINTER(C.c_int)]
    L.pt_bvh_build_sah.argtypes = [C.c_int32] + [C.c_void_p] * 12
    _libs[numerics] = L
    return L


def check(rc: int, what: str = "", L: C.CDLL | None = None):
    if rc != 0:
        msg = (L or lib()).pt_last_error()       # the error slot lives in the library that failed
        raise B200PTError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


_MATERIAL_KEYS = (("material_type", "i", ()), ("material_albedo", "f", (3,)), ("material_fuzz", "f", ()),
                  ("material_ir", "f", ()), ("material_emit_color", "f", (3,)), ("texture_type", "i", ()),
                  ("texture_scale", "f", ()), ("texture_color1", "f", (3,)), ("texture_color2", "f", (3,)),
                  ("texture_image_idx", "i", ()), ("is_constant_medium", "i", ()), ("medium_density", "f", ()),
                  ("medium_albedo", "f", (3,)))
_MATERIAL_FIELDS = ("material_type", "albedo", "fuzz", "ir", "emit", "texture_type", "texture_scale", "color1", "color2",
                    "image_idx", "is_medium", "medium_density", "medium_albedo")


class HostScene:
    """Contiguous typed copies of a CompiledScene dict + the PtHostScene view over them.

    The dict is the flat form scene_compiler.compile_scene + bvh.compile_bvh produce (the reference's
    compile_scene returns the same arrays as a tuple of per-type dicts, scene_compiler.py:931-965).
    A primitive type that is present must come with ALL of its arrays, each with one row per
    primitive: a missing key or a short array raises here instead of handing the C packer a NULL or
    an out-of-bounds pointer (the packer re-checks pointers and index ranges: PT_ERR_INVALID_ARG)."""

    def __init__(self, cs: dict):
        self._keep = []
        s = PtHostScene()

        def arr(key, kind, tail, rows=None):
            if key not in cs:
                if rows:
                    raise B200PTError(f"compiled scene lacks '{key}' ({rows} primitives of that type need it)")
                a = np.zeros((0,) + tail, np.float32 if kind == "f" else np.int32)
            else:
                a = np.ascontiguousarray(cs[key], dtype=np.float32 if kind == "f" else np.int32)
                if a.shape[1:] != tail and not (tail == () and a.ndim == 1):
                    raise B200PTError(f"compiled scene: '{key}' has shape {a.shape}, expected (n,{','.join(map(str, tail))})")
                if rows is not None and a.shape[0] != rows:
                    raise B200PTError(f"compiled scene: '{key}' has {a.shape[0]} rows for {rows} primitives")
            self._keep.append(a)
            return a

        sd = arr("sphere_data", "f", (4,))
        nq = len(cs["quad_Q"]) if "quad_Q" in cs else 0
        nt = len(cs["triangle_v0"]) if "triangle_v0" in cs else 0
        s.num_spheres, s.sphere_data = sd.shape[0], _ptr(sd)
        s.num_quads, s.num_triangles = nq, nt
        for k in ("quad_Q", "quad_u", "quad_v", "quad_normal", "quad_D", "quad_w"):
            setattr(s, k, _ptr(arr(k, "f", () if k == "quad_D" else (3,), nq)))
        for k, dst in (("triangle_v0", "tri_v0"), ("triangle_edge1", "tri_edge1"), ("triangle_edge2", "tri_edge2"),
                       ("triangle_normal", "tri_normal")):
            setattr(s, dst, _ptr(arr(k, "f", (3,), nt)))
        counts = {PRIM_SPHERE: sd.shape[0], PRIM_TRIANGLE: nt, PRIM_QUAD: nq}
        for pt, prefix in PRIM_PREFIX.items():
            m = s.mat[pt]
            for (key, kind, tail), field in zip(_MATERIAL_KEYS, _MATERIAL_FIELDS):
                setattr(m, field, _ptr(arr(f"{prefix}_{key}", kind, tail, counts[pt])))
        bmin = arr("bvh_bbox_min", "f", (3,))
        nn = bmin.shape[0]
        if nn == 0 and sum(counts.values()) > 0:
            raise B200PTError("compiled scene has primitives but no BVH arrays (run bvh.compile_bvh)")
        s.num_nodes, s.bvh_bbox_min = nn, _ptr(bmin)
        s.bvh_bbox_max = _ptr(arr("bvh_bbox_max", "f", (3,), nn))
        s.bvh_left = _ptr(arr("bvh_left_child", "i", (), nn))
        s.bvh_right = _ptr(arr("bvh_right_child", "i", (), nn))
        s.bvh_prim_type = _ptr(arr("bvh_prim_type", "i", (), nn))
        s.bvh_prim_idx = _ptr(arr("bvh_prim_idx", "i", (), nn))
        rv = np.ascontiguousarray(cs["perlin_randvec"] if "perlin_randvec" in cs else np.zeros((256, 3)), np.float32)
        if rv.shape != (256, 3):
            raise B200PTError(f"compiled scene: perlin_randvec has shape {rv.shape}, expected (256, 3)")
        self._keep.append(rv)
        s.perlin_randvec = _ptr(rv)
        for k, dst in (("perlin_perm_x", "perm_x"), ("perlin_perm_y", "perm_y"), ("perlin_perm_z", "perm_z")):
            a = np.ascontiguousarray(cs[k] if k in cs else np.arange(256), np.int32)
            if a.shape != (256,):
                raise B200PTError(f"compiled scene: {k} has shape {a.shape}, expected (256,)")
            self._keep.append(a)
            setattr(s, dst, _ptr(a))
        n_img = int(cs.get("num_images", 0))
        if n_img > PT_MAX_IMAGES:
            raise B200PTError(f"{n_img} image textures > {PT_MAX_IMAGES}")
        s.num_images = n_img
        for i in range(n_img):
            if f"image_{i}" not in cs:
                raise B200PTError(f"compiled scene lacks 'image_{i}' (num_images = {n_img})")
            img = np.ascontiguousarray(cs[f"image_{i}"], dtype=np.uint8)
            if img.ndim != 3 or img.shape[2] != 3:
                raise B200PTError(f"compiled scene: image_{i} has shape {img.shape}, expected (h, w, 3) uint8")
            self._keep.append(img)
            s.images[i] = img.ctypes.data
            s.image_h[i], s.image_w[i] = img.shape[0], img.shape[1]
        s.full_features = int(cs.get("full_features", 0))
        if s.full_features:
            mo = np.ascontiguousarray(cs["sphere_motion"] if "sphere_motion" in cs else np.zeros((sd.shape[0], 3)), np.float32)
            if mo.shape != (sd.shape[0], 3):
                raise B200PTError(f"compiled scene: sphere_motion has shape {mo.shape} for {sd.shape[0]} spheres")
            self._keep.append(mo)
            s.sphere_motion = _ptr(mo)
        self.c = s
        self.input_bytes = int(sum(a.nbytes for a in self._keep))


JITTER = {"random": 0, "sobol": 1}


def make_camera(cam: dict, background=(0.7, 0.8, 1.0), max_depth=50, jitter: str = "random") -> PtCamera:
    """The float64 -> float32 narrowing of _upload_camera (reference renderer.py:230-247)."""
    c = PtCamera()
    for src, dst in (("cam_center", "center"), ("cam_pixel00", "pixel00"), ("cam_delta_u", "delta_u"),
                     ("cam_delta_v", "delta_v"), ("cam_defocus_disk_u", "defocus_disk_u"),
                     ("cam_defocus_disk_v", "defocus_disk_v")):
        v = np.asarray(cam[src], dtype=np.float64).astype(np.float32)
        getattr(c, dst)[:] = [float(x) for x in v]
    c.defocus_angle = float(np.float32(cam["cam_defocus_angle"]))
    c.background[:] = [float(np.float32(x)) for x in background]
    c.max_depth = int(max_depth)
    if jitter not in JITTER:
        raise B200PTError(f"jitter must be one of {sorted(JITTER)}")
    c.jitter = JITTER[jitter]
    return c
