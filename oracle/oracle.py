"""TEST INFRASTRUCTURE ONLY: ctypes binding of the CPU parity oracle (oracle/oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (path_tracer_python_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

PRIM_PREFIX = {0: "sphere", 1: "triangle", 2: "quad"}  # prim_type -> array prefix


class OMaterials(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "material_type", "albedo", "fuzz", "ir", "emit", "texture_type", "texture_scale", "color1",
        "color2", "image_idx", "is_medium", "medium_density", "medium_albedo")]


class OScene(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in ("num_spheres", "num_quads", "num_triangles", "num_nodes", "num_images")]
        + [(n, C.c_void_p) for n in ("sphere_data", "quad_Q", "quad_u", "quad_v", "quad_normal", "quad_D",
                                     "quad_w", "tri_v0", "tri_edge1", "tri_edge2", "tri_normal")]
        + [("mat", OMaterials * 3)]
        + [(n, C.c_void_p) for n in ("bvh_bbox_min", "bvh_bbox_max", "bvh_left", "bvh_right", "bvh_prim_type",
                                     "bvh_prim_idx", "perlin_randvec", "perm_x", "perm_y", "perm_z")]
        + [("images", C.c_void_p * 16), ("image_w", C.c_int32 * 16), ("image_h", C.c_int32 * 16)]
        + [("sphere_motion", C.c_void_p), ("full_features", C.c_int32)]
    )


class OCamera(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("pixel00", C.c_float * 3), ("delta_u", C.c_float * 3),
                ("delta_v", C.c_float * 3), ("disk_u", C.c_float * 3), ("disk_v", C.c_float * 3),
                ("defocus_angle", C.c_float), ("bg", C.c_float * 3), ("max_depth", C.c_int32), ("jitter", C.c_int32)]


class OCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "node_tests", "sphere_tests", "triangle_tests", "quad_tests",
                                          "paths", "bounces")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(force: bool = False) -> str:
    """Compile oracle/oracle.cpp -> oracle/_build/liboracle.so (g++, strict IEEE flags)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "oracle_math.h")]
    if not force and os.path.exists(_LIB_PATH) and all(
            os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    base = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
            "-fno-unsafe-math-optimizations", "-o", _LIB_PATH, srcs[0]]
    r = subprocess.run(base[:1] + ["-fopenmp"] + base[1:], capture_output=True, text=True)
    if r.returncode != 0:  # no libgomp: single-threaded oracle
        r = subprocess.run(base, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stderr)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.o_rng_seed_state.restype = C.c_uint64
        _lib.o_rng_seed_state.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.o_num_threads.restype = C.c_int
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class OracleScene:
    """Holds contiguous copies of the reference-layout arrays and the ctypes view of them."""

    def __init__(self, cs: dict):
        self._keep = []
        s = OScene()

        def f32(key, shape_tail=()):
            a = np.ascontiguousarray(cs.get(key, np.zeros((0,) + shape_tail)), dtype=np.float32)
            self._keep.append(a)
            return a

        def i32(key):
            a = np.ascontiguousarray(cs.get(key, np.zeros((0,))), dtype=np.int32)
            self._keep.append(a)
            return a

        sd = f32("sphere_data", (4,))
        s.num_spheres = sd.shape[0]
        s.sphere_data = _ptr(sd)
        for k in ("quad_Q", "quad_u", "quad_v", "quad_normal", "quad_D", "quad_w"):
            a = f32(k, () if k == "quad_D" else (3,))
            setattr(s, k, _ptr(a))
            if k == "quad_Q":
                s.num_quads = a.shape[0]
        for k, dst in (("triangle_v0", "tri_v0"), ("triangle_edge1", "tri_edge1"),
                       ("triangle_edge2", "tri_edge2"), ("triangle_normal", "tri_normal")):
            a = f32(k, (3,))
            setattr(s, dst, _ptr(a))
            if dst == "tri_v0":
                s.num_triangles = a.shape[0]
        for pt, prefix in PRIM_PREFIX.items():
            m = s.mat[pt]
            m.material_type = _ptr(i32(f"{prefix}_material_type"))
            m.albedo = _ptr(f32(f"{prefix}_material_albedo", (3,)))
            m.fuzz = _ptr(f32(f"{prefix}_material_fuzz"))
            m.ir = _ptr(f32(f"{prefix}_material_ir"))
            m.emit = _ptr(f32(f"{prefix}_material_emit_color", (3,)))
            m.texture_type = _ptr(i32(f"{prefix}_texture_type"))
            m.texture_scale = _ptr(f32(f"{prefix}_texture_scale"))
            m.color1 = _ptr(f32(f"{prefix}_texture_color1", (3,)))
            m.color2 = _ptr(f32(f"{prefix}_texture_color2", (3,)))
            m.image_idx = _ptr(i32(f"{prefix}_texture_image_idx"))
            m.is_medium = _ptr(i32(f"{prefix}_is_constant_medium"))
            m.medium_density = _ptr(f32(f"{prefix}_medium_density"))
            m.medium_albedo = _ptr(f32(f"{prefix}_medium_albedo", (3,)))
        bmin = f32("bvh_bbox_min", (3,))
        s.num_nodes = bmin.shape[0]
        s.bvh_bbox_min = _ptr(bmin)
        s.bvh_bbox_max = _ptr(f32("bvh_bbox_max", (3,)))
        s.bvh_left = _ptr(i32("bvh_left_child"))
        s.bvh_right = _ptr(i32("bvh_right_child"))
        s.bvh_prim_type = _ptr(i32("bvh_prim_type"))
        s.bvh_prim_idx = _ptr(i32("bvh_prim_idx"))
        rv = cs.get("perlin_randvec")
        if rv is None:
            rv = np.zeros((256, 3), np.float32)
        rv = np.ascontiguousarray(rv, dtype=np.float32)
        self._keep.append(rv)
        s.perlin_randvec = _ptr(rv)
        for k, dst in (("perlin_perm_x", "perm_x"), ("perlin_perm_y", "perm_y"), ("perlin_perm_z", "perm_z")):
            a = cs.get(k)
            if a is None:
                a = np.arange(256)
            a = np.ascontiguousarray(a, dtype=np.int32)
            self._keep.append(a)
            setattr(s, dst, _ptr(a))
        n_img = int(cs.get("num_images", 0))
        s.num_images = n_img
        for i in range(n_img):
            img = np.ascontiguousarray(cs[f"image_{i}"], dtype=np.uint8)
            self._keep.append(img)
            s.images[i] = img.ctypes.data
            s.image_h[i], s.image_w[i] = img.shape[0], img.shape[1]
        s.full_features = int(cs.get("full_features", 0))
        if s.full_features and "sphere_motion" in cs:
            s.sphere_motion = _ptr(f32("sphere_motion", (3,)))
        self.c = s


def make_camera(cam: dict, bg=(0.7, 0.8, 1.0), max_depth=50, jitter: str = "random") -> OCamera:
    """cam: dict with cam_center/... (float64, as uploaded by renderer.py:230-247 -> f32)."""
    c = OCamera()
    for src, dst in (("cam_center", "center"), ("cam_pixel00", "pixel00"), ("cam_delta_u", "delta_u"),
                     ("cam_delta_v", "delta_v"), ("cam_defocus_disk_u", "disk_u"),
                     ("cam_defocus_disk_v", "disk_v")):
        v = np.asarray(cam[src], dtype=np.float64).astype(np.float32)
        getattr(c, dst)[:] = [float(x) for x in v]
    c.defocus_angle = float(np.float32(cam["cam_defocus_angle"]))
    c.bg[:] = [float(np.float32(x)) for x in bg]
    c.max_depth = int(max_depth)
    c.jitter = {"random": 0, "sobol": 1}[jitter]
    return c


def render(scene: OracleScene, cam: OCamera, W, H, spp, seed=0, mode="megakernel", accum=None, y0=0, y1=None,
           sample_begin=0, nthreads=0):
    """Returns (accum HxWx3 float32, counters dict)."""
    if accum is None:
        accum = np.zeros((H, W, 3), dtype=np.float32)
    assert accum.dtype == np.float32 and accum.flags.c_contiguous
    y1 = H if y1 is None else y1
    cnt = OCounters()
    fn = lib().o_render_megakernel if mode == "megakernel" else lib().o_render_wavefront
    fn(C.byref(scene.c), C.byref(cam), C.c_int(W), C.c_int(H), C.c_int(y0), C.c_int(y1), C.c_int(sample_begin),
       C.c_int(spp), C.c_uint32(seed), accum.ctypes.data_as(C.c_void_p), C.byref(cnt), C.c_int(nthreads))
    return accum, cnt.as_dict()


def generate_camera_rays(cam: OCamera, W, H, seed=0, sample=0):
    o = np.zeros((H * W, 3), np.float32)
    d = np.zeros((H * W, 3), np.float32)
    rng = np.zeros((H * W,), np.uint32)
    lib().o_generate_camera_rays(C.byref(cam), C.c_int(W), C.c_int(H), C.c_uint32(seed), C.c_int(sample),
                                 o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                                 rng.ctypes.data_as(C.c_void_p))
    return o, d, rng


def first_hit(scene: OracleScene, o, d, t_min=0.001, t_max=1e10, nthreads=0):
    o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
    n = o.shape[0]
    out = {"valid": np.zeros(n, np.int32), "t": np.zeros(n, np.float32), "prim_type": np.zeros(n, np.int32),
           "prim_idx": np.zeros(n, np.int32), "p": np.zeros((n, 3), np.float32), "n": np.zeros((n, 3), np.float32)}
    cnt = OCounters()
    lib().o_first_hit(C.byref(scene.c), C.c_int(n), o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                      C.c_float(t_min), C.c_float(t_max), *[out[k].ctypes.data_as(C.c_void_p) for k in
                                                             ("valid", "t", "prim_type", "prim_idx", "p", "n")],
                      C.byref(cnt), C.c_int(nthreads))
    out["counters"] = cnt.as_dict()
    return out


_FN = {"sin": 0, "cos": 1, "log": 2, "acos": 3, "atan2": 4, "pow5": 5, "exp": 7}


def math_eval(fn: str, a, b=None):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b if b is not None else np.zeros_like(a), np.float32)
    out = np.empty_like(a)
    lib().o_math_eval(C.c_int(_FN[fn]), C.c_int(a.size), a.ctypes.data_as(C.c_void_p),
                      b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def rng_seed(seed, pixel, sample) -> int:
    """Initial state of the path's random stream as one 64-bit word: stride << 32 | position."""
    return int(lib().o_rng_seed_state(C.c_uint32(seed), C.c_uint32(pixel), C.c_uint32(sample)))


def rng_draw(state: int, n: int):
    st = C.c_uint64(state)
    out = np.empty(n, np.float32)
    lib().o_rng_draw(C.byref(st), C.c_int(n), out.ctypes.data_as(C.c_void_p))
    return out, int(st.value)


def ld_jitter_points(seed, pixel, n):
    """The first n points of a pixel's scrambled Sobol' (0,2)-sequence: (n, 2) float32 in [0, 1)."""
    out = np.empty((n, 2), np.float32)
    lib().o_ld_jitter_points(C.c_uint32(seed), C.c_uint32(pixel), C.c_int(n), out.ctypes.data_as(C.c_void_p))
    return out


def rng_seed_states(seed, pixels, samples):
    """Vectorised rng_seed: uint64 stream states of the (pixel, sample) keys."""
    px = np.ascontiguousarray(pixels, np.uint32)
    sm = np.ascontiguousarray(samples, np.uint32)
    out = np.empty(px.size, np.uint64)
    lib().o_rng_seed_states(C.c_uint32(seed), C.c_int64(px.size), px.ctypes.data_as(C.c_void_p),
                            sm.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def rng_first_draws(states, k):
    """The first k raw 32-bit outputs of each stream: (n, k) uint32."""
    st = np.ascontiguousarray(states, np.uint64)
    out = np.empty((st.size, k), np.uint32)
    lib().o_rng_first_draws(C.c_int64(st.size), st.ctypes.data_as(C.c_void_p), C.c_int(k), out.ctypes.data_as(C.c_void_p))
    return out


def buffer_to_image(accum, sample_count):
    a = np.ascontiguousarray(accum, np.float32)
    out = np.empty(a.shape, np.uint8)
    lib().o_buffer_to_image(a.ctypes.data_as(C.c_void_p), C.c_int(a.size), C.c_int(sample_count),
                            out.ctypes.data_as(C.c_void_p))
    return out


def num_threads() -> int:
    return int(lib().o_num_threads())
