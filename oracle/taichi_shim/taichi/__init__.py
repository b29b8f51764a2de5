"""TEST INFRASTRUCTURE ONLY: a pure-Python stand-in for the `taichi` module.

`taichi` is the reference's un-vendored, un-pinned JIT dependency and is absent from this image
(SURVEY.md section 8c).  This shim implements just the API subset that
/root/reference/src/render_server/taichi_renderer/{fields,kernels,renderer}.py use, with Taichi's
python-scope *and* kernel-scope semantics collapsed into ordinary Python execution, so that the
reference's own, unmodified device code can be executed here to mint golden vectors
(oracle/gen_golden.py).  It is slow (microseconds per scalar operation) and only meant for tiny
images.

Numerics = the PT-F32 contract (oracle/oracle_math.h):
  * every f32 value is a numpy.float32 scalar, so each source-level operation rounds once to
    binary32 (Taichi default_fp=f32 with fast_math disabled and no FMA contraction);
  * `ti.sin/cos/log/acos/atan2/pow` are the contract's functions (evaluated by liboracle.so's
    o_math_eval so that shim and oracle share one definition of the *intrinsics*, while all
    control flow and arithmetic of kernels.py is executed from the reference's source);
  * `ti.random()` draws from the contract's per-path stream keyed by (seed, pixel, sample).  The
    "current thread" is the pixel whose struct-for iteration is running (megakernel,
    generate_camera_rays) or whose `ray_pixel_index[i]` was read last (wavefront stages); the
    sample index is the number of `render_sample` / `generate_camera_rays` launches since the last
    `clear_accum_buffer`.
"""
import itertools
import os
import sys

import numpy as _np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as _oracle  # noqa: E402  (oracle/oracle.py)

_F = _np.float32


# ----------------------------------------------------------------------------- dtypes / init
class _DType:
    def __init__(self, name, np_type, is_float):
        self.name, self.np_type, self.is_float = name, np_type, is_float

    def __repr__(self):
        return f"ti.{self.name}"


f32 = _DType("f32", _np.float32, True)
i32 = _DType("i32", _np.int32, False)
u8 = _DType("u8", _np.uint8, False)
metal = "metal"
cpu = "cpu"
gpu = "gpu"
cuda = "cuda"


def init(*args, **kwargs):
    return None


def sync():
    return None


def static(x):
    return x


def ndrange(*dims):
    return itertools.product(*[range(int(d)) for d in dims])


# ----------------------------------------------------------------------------- RNG / thread context
class _Runtime:
    def __init__(self):
        self.seed = 0
        self.sample_counter = 0      # launches since last clear
        self.current_sample = 0
        self.current_pixel = None
        self.states = {}             # pixel -> u64 stream state (stride << 32 | position) for current sample
        self.image_width = None      # set by the driver (needed to turn (py,px) into a pixel index)
        self.draws = 0

    def begin_sample(self):
        self.current_sample = self.sample_counter
        self.sample_counter += 1
        self.states = {}

    def set_pixel(self, pixel_index):
        self.current_pixel = int(pixel_index)

    def next_f32(self):
        p = self.current_pixel
        s = self.states.get(p)
        if s is None:
            s = _oracle.rng_seed(self.seed, p, self.current_sample)      # stride << 32 | position
        s = (s & 0xFFFFFFFF00000000) | ((s + (s >> 32)) & 0xFFFFFFFF)   # position += stride
        self.states[p] = s
        z = s & 0xFFFFFFFF
        z ^= z >> 16
        z = (z * 0x21F0AAAD) & 0xFFFFFFFF
        z ^= z >> 15
        z = (z * 0x735A2D97) & 0xFFFFFFFF
        z ^= z >> 15
        self.draws += 1
        return _F(z >> 8) * _F(5.9604644775390625e-08)


runtime = _Runtime()


def random(dt=f32):
    return runtime.next_f32()


# kernels whose launch advances / resets the sample counter
_SAMPLE_START_KERNELS = {"render_sample", "generate_camera_rays"}
_SAMPLE_RESET_KERNELS = {"clear_accum_buffer", "clear_accumulation_buffer"}


def _coerce(value, ann):
    if ann is f32:
        return _F(value)
    if ann is i32:
        return int(value)
    return value


def _wrap(fn, is_kernel):
    import functools
    import inspect

    sig = inspect.signature(fn)
    params = list(sig.parameters.values())
    anns = [p.annotation for p in params]

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if is_kernel:
            if fn.__name__ in _SAMPLE_RESET_KERNELS:
                runtime.sample_counter = 0
            if fn.__name__ in _SAMPLE_START_KERNELS:
                runtime.begin_sample()
        args = tuple(_coerce(a, anns[i]) if i < len(anns) else a for i, a in enumerate(args))
        if fn.__name__ == "get_ray":
            # get_ray(px, py) is the first thing a path's thread does (kernels.py:1185, :1229);
            # generate_camera_rays iterates a plain ndrange, so the thread context is set here
            runtime.set_pixel(args[1] * runtime.image_width + args[0])
        return fn(*args, **kwargs)

    return wrapper


def func(fn):
    return _wrap(fn, False)


def kernel(fn):
    return _wrap(fn, True)


# ----------------------------------------------------------------------------- scalar helpers
def _is_vec(x):
    return isinstance(x, Vec)


def _f(x):
    return x if isinstance(x, _np.float32) else _F(x)


def cast(x, dt):
    if _is_vec(x):
        return Vec([cast(e, dt) for e in x.e])
    if dt.is_float:
        return _F(x)
    return int(x)  # truncation toward zero, like a C cast


def _m1(fn_name, x):
    return _oracle.math_eval(fn_name, _np.array([x], _np.float32))[0]


def sqrt(x):
    return _np.sqrt(_f(x))


def sin(x):
    return _m1("sin", x)


def cos(x):
    return _m1("cos", x)


def log(x):
    return _m1("log", x)


def acos(x):
    return _m1("acos", x)


def atan2(y, x):
    return _oracle.math_eval("atan2", _np.array([y], _np.float32), _np.array([x], _np.float32))[0]


def pow(x, y):  # noqa: A001
    assert float(y) == 5.0, "contract defines pow only for the literal exponent 5 (kernels.py:786)"
    return _m1("pow5", x)


def floor(x):
    return _np.floor(_f(x))


def abs(x):  # noqa: A001
    if isinstance(x, int):
        return -x if x < 0 else x
    return _np.abs(_f(x))


def _minmax(a, b, is_min):
    if _is_vec(a) or _is_vec(b):
        ae = a.e if _is_vec(a) else [a] * len(b.e)
        be = b.e if _is_vec(b) else [b] * len(a.e)
        return Vec([_minmax(x, y, is_min) for x, y in zip(ae, be)])
    if isinstance(a, int) and isinstance(b, int):
        return (a if a < b else b) if is_min else (a if a > b else b)
    a, b = _f(a), _f(b)
    if is_min:
        return a if a < b else b
    return a if a > b else b


def min(a, b):  # noqa: A001
    return _minmax(a, b, True)


def max(a, b):  # noqa: A001
    return _minmax(a, b, False)


# ----------------------------------------------------------------------------- vectors
class Vec:
    """ti.math.vec3 / ti.Vector value.  Immutable arithmetic; __setitem__ only for local stacks."""
    __slots__ = ("e", "_ref")
    __array_ufunc__ = None  # numpy scalars must defer to __radd__/__rmul__/... instead of broadcasting

    def __init__(self, elems, dt=None):
        if dt is not None and not dt.is_float:
            self.e = [int(x) for x in elems]
        elif dt is not None:
            self.e = [_F(x) for x in elems]
        else:
            self.e = list(elems)
        self._ref = None

    # component access
    @property
    def x(self):
        return self.e[0]

    @property
    def y(self):
        return self.e[1]

    @property
    def z(self):
        return self.e[2]

    def __getitem__(self, i):
        return self.e[int(i)]

    def __setitem__(self, i, v):
        self.e[int(i)] = v

    def __len__(self):
        return len(self.e)

    def __iter__(self):
        return iter(self.e)

    def _bin(self, other, op):
        if _is_vec(other):
            return Vec([op(a, b) for a, b in zip(self.e, other.e)])
        return Vec([op(a, other) for a in self.e])

    def _rbin(self, other, op):
        return Vec([op(other, a) for a in self.e])

    def __add__(self, o):
        return self._bin(o, lambda a, b: a + b)

    def __radd__(self, o):
        return self._rbin(o, lambda a, b: a + b)

    def __sub__(self, o):
        return self._bin(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return self._rbin(o, lambda a, b: a - b)

    def __mul__(self, o):
        return self._bin(o, lambda a, b: a * b)

    def __rmul__(self, o):
        return self._rbin(o, lambda a, b: a * b)

    def __truediv__(self, o):
        return self._bin(o, lambda a, b: a / b)

    def __neg__(self):
        return Vec([-a for a in self.e])

    def dot(self, o):
        acc = self.e[0] * o.e[0]
        for a, b in zip(self.e[1:], o.e[1:]):
            acc = acc + a * b
        return acc

    def cross(self, o):
        a, b = self.e, o.e
        return Vec([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])

    def norm_sqr(self):
        return self.dot(self)

    def norm(self):
        return sqrt(self.norm_sqr())

    def normalized(self, eps=0):
        invlen = 1 / (self.norm() + eps)
        return invlen * self

    def to_list(self):
        return list(self.e)

    def __repr__(self):
        return f"Vec({self.e})"


def _vec_ctor(n):
    def ctor(*args):
        if len(args) == 1 and not isinstance(args[0], (list, tuple, Vec, _np.ndarray)):
            return Vec([_F(args[0])] * n)
        if len(args) == 1:
            args = tuple(args[0])
        assert len(args) == n
        return Vec([_F(a) for a in args])

    return ctor


# ----------------------------------------------------------------------------- fields
class _IntRef(int):
    """Value of a 0-d i32 field that remembers where it came from (for ti.atomic_add)."""
    _field = None


class ScalarField:
    def __init__(self, dt, shape):
        if shape is None:
            shape = ()
        if isinstance(shape, int):
            shape = (shape,)
        self.dt = dt
        self.shape = tuple(shape)
        self.a = _np.zeros(self.shape, dtype=dt.np_type)
        self.on_read = None

    def _idx(self, i):
        if i is None:
            return ()
        return i

    def __getitem__(self, i):
        v = self.a[self._idx(i)]
        if self.on_read is not None:
            self.on_read(v)
        if self.dt.is_float:
            return _F(v)
        if self.shape == ():
            r = _IntRef(int(v))
            r._field = self
            return r
        return int(v)

    def __setitem__(self, i, v):
        self.a[self._idx(i)] = v

    def __iter__(self):
        for idx in _np.ndindex(*self.shape):
            yield idx if len(idx) > 1 else idx[0]

    def from_numpy(self, arr):
        self.a[...] = arr

    def to_numpy(self):
        return self.a.copy()

    def fill(self, v):
        self.a[...] = v


class VectorField:
    def __init__(self, n, dt, shape):
        if shape is None:
            shape = ()
        if isinstance(shape, int):
            shape = (shape,)
        self.n, self.dt = n, dt
        self.shape = tuple(shape)
        self.a = _np.zeros(self.shape + (n,), dtype=dt.np_type)
        self.struct_for_hook = None

    def _idx(self, i):
        if i is None:
            return ()
        return i

    def __getitem__(self, i):
        row = self.a[self._idx(i)]
        if self.dt.is_float:
            v = Vec([_F(x) for x in row])
        else:
            v = Vec([int(x) for x in row])
        v._ref = (self, self._idx(i))
        return v

    def __setitem__(self, i, v):
        if _is_vec(v):
            v = v.e
        self.a[self._idx(i)] = _np.asarray(v, dtype=self.dt.np_type)

    def __iter__(self):
        # struct-for: one "GPU thread" per element; the 2-D accum buffer is H x W
        for idx in _np.ndindex(*self.shape):
            if len(idx) == 2:
                runtime.set_pixel(idx[0] * self.shape[1] + idx[1])
            yield idx if len(idx) > 1 else idx[0]

    def from_numpy(self, arr):
        self.a[...] = arr

    def to_numpy(self):
        return self.a.copy()

    def fill(self, v):
        self.a[...] = v


def field(dt, shape=None):
    return ScalarField(dt, shape)


class _VectorNS:
    """ti.Vector([...], dt=...) and ti.Vector.field(n, dt, shape)."""

    def __call__(self, elems, dt=None):
        return Vec(elems, dt=dt)

    @staticmethod
    def field(n, dtype, shape=None):
        return VectorField(n, dtype, shape)


Vector = _VectorNS()


def atomic_add(target, value):
    if isinstance(target, _IntRef):
        f = target._field
        old = int(f.a[()])
        f.a[()] = old + int(value)
        return old
    if _is_vec(target) and target._ref is not None:
        f, idx = target._ref
        old = f[idx]
        f[idx] = old + value
        return old
    raise TypeError("atomic_add target is not a field element")


# ----------------------------------------------------------------------------- ti.dataclass
class _StructProxy:
    __slots__ = ("_sf", "_i")

    def __init__(self, sf, i):
        object.__setattr__(self, "_sf", sf)
        object.__setattr__(self, "_i", i)

    def __getattr__(self, name):
        return self._sf.members[name][self._i]

    def __setattr__(self, name, value):
        self._sf.members[name][self._i] = value


class _StructField:
    def __init__(self, annotations, shape):
        self.members = {}
        for name, ann in annotations.items():
            if ann is math.vec3:
                self.members[name] = VectorField(3, f32, shape)
            elif ann is math.vec4:
                self.members[name] = VectorField(4, f32, shape)
            elif isinstance(ann, _DType):
                self.members[name] = ScalarField(ann, shape)
            else:
                raise TypeError(f"unsupported struct member {name}: {ann}")

    def __getitem__(self, i):
        return _StructProxy(self, i)


def dataclass(cls):
    anns = dict(cls.__annotations__)

    class _Struct:
        @staticmethod
        def field(shape=None):
            return _StructField(anns, shape)

    _Struct.__name__ = cls.__name__
    return _Struct


# ----------------------------------------------------------------------------- ti.math
class _MathNS:
    pass


math = _MathNS()
math.vec2 = _vec_ctor(2)
math.vec3 = _vec_ctor(3)
math.vec4 = _vec_ctor(4)
math.pi = 3.141592653589793
sys.modules[__name__ + ".math"] = math  # `import taichi.math` / `ti.math.vec3`


class _Profiler:
    @staticmethod
    def print_kernel_profiler_info(*a, **k):
        return None

    @staticmethod
    def clear_kernel_profiler_info(*a, **k):
        return None


profiler = _Profiler()
