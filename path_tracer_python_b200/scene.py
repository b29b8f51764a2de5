"""Light-weight scene description classes with the reference's names and attributes
(src/core/{sphere,quad,triangle,hittable_list,constant_medium,material,texture,camera}.py), for
building scenes where the reference's `src/core` package is not importable (the GPU boxes, the
benchmarks, the tests).  They carry data only -- no CPU ``hit()`` / ``scatter()``: tracing happens
on the GPU.  ``scene_compiler.compile_scene`` accepts these and the reference's own objects alike.
"""
from __future__ import annotations

import math

import numpy as np


class vec3:
    __slots__ = ("x", "y", "z")

    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def __add__(self, o):
        return vec3(self.x + o.x, self.y + o.y, self.z + o.z)

    def __sub__(self, o):
        return vec3(self.x - o.x, self.y - o.y, self.z - o.z)

    def __neg__(self):
        return vec3(-self.x, -self.y, -self.z)

    def __mul__(self, s):
        if isinstance(s, vec3):
            return vec3(self.x * s.x, self.y * s.y, self.z * s.z)
        return vec3(self.x * s, self.y * s, self.z * s)

    __rmul__ = __mul__

    def __truediv__(self, s):
        return vec3(self.x / s, self.y / s, self.z / s)

    def dot(self, o):
        return self.x * o.x + self.y * o.y + self.z * o.z

    def cross(self, o):
        return vec3(self.y * o.z - self.z * o.y, self.z * o.x - self.x * o.z, self.x * o.y - self.y * o.x)

    def length(self):
        return math.sqrt(self.x ** 2 + self.y ** 2 + self.z ** 2)

    def unit_vector(self):
        return self / self.length()

    def __repr__(self):
        return f"vec3({self.x}, {self.y}, {self.z})"


point3 = vec3
color = vec3


class _Path:
    """center(time) = origin + time * direction, like the reference's Ray-valued Sphere.center."""

    def __init__(self, origin, direction):
        self.origin, self.direction = origin, direction

    def at(self, t):
        return self.origin + self.direction * t


# ---- textures
class solid_color:
    def __init__(self, albedo):
        self.albedo = albedo

    @classmethod
    def from_color(cls, albedo):
        return cls(albedo)

    def value(self, u, v, p):
        return self.albedo


class checker_texture:
    def __init__(self, scale, even, odd):
        self.inv_scale = 1.0 / scale
        self.even, self.odd = even, odd

    @classmethod
    def from_colors(cls, scale, c1, c2):
        return cls(scale, solid_color(c1), solid_color(c2))


class noise_texture:
    def __init__(self, scale=1.0):
        # the reference's noise_texture owns a perlin() (core/texture.py:80-84), whose constructor draws its
        # gradient vectors and permutations from Python's `random`: the same draws are made here so that a
        # seeded scene script continues with the same stream (the Taichi path itself uses the renderer's tables)
        from .scene_compiler import perlin_tables
        self.noise = perlin_tables()
        self.scale = scale


class _Image:
    def __init__(self, u8):
        self.bdata8 = np.ascontiguousarray(u8, np.uint8)
        self.fdata = self.bdata8.astype(np.float32) / np.float32(255.0)


class image_texture:
    """8-bit RGB image; `source` is a file name (PIL) or an (H, W, 3) uint8 array."""

    def __init__(self, source):
        if isinstance(source, np.ndarray):
            self.image = _Image(source)
        else:
            from PIL import Image
            self.image = _Image(np.array(Image.open(source).convert("RGB"), dtype=np.uint8))


# ---- materials
class lambertian:
    def __init__(self, tex):
        self.tex = tex

    @classmethod
    def from_color(cls, albedo):
        return cls(solid_color(albedo))

    @classmethod
    def from_texture(cls, tex):
        return cls(tex)


class metal:
    def __init__(self, albedo, fuzz):
        self.albedo = albedo
        self.fuzz = fuzz if fuzz < 1.0 else 1.0


class dielectric:
    def __init__(self, index_of_refraction):
        self.ir = index_of_refraction


class diffuse_light:
    def __init__(self, tex):
        self.tex = tex

    @classmethod
    def from_color(cls, emit_color):
        return cls(solid_color(emit_color))


class isotropic:
    def __init__(self, tex):
        self.tex = tex

    @classmethod
    def from_color(cls, albedo):
        return cls(solid_color(albedo))


class subsurface_simple:
    """core/material.py:144-173.  Rendered by the `full` feature mode; the reference's Taichi path
    (and this renderer's default `taichi_compat` mode) turns it into 0.8-grey Lambertian."""

    def __init__(self, albedo, scatter_distance):
        self.albedo, self.scatter_distance = albedo, scatter_distance


class subsurface_volumetric:
    """core/material.py:175-276 (random walk with Henyey-Greenstein phase function)."""

    def __init__(self, albedo, scatter_coeff=0.08, absorb_coeff=0.8, g=0.0):
        self.albedo = albedo
        self.sigma_s, self.sigma_a, self.g = scatter_coeff, absorb_coeff, g
        self.sigma_t = scatter_coeff + absorb_coeff


# ---- geometry
class Sphere:
    def __init__(self, center_path, radius, mat):
        self.center = center_path
        self.radius = max(0.0, radius)
        self.material = mat

    @classmethod
    def stationary(cls, static_center, radius, mat):
        return cls(_Path(static_center, vec3(0, 0, 0)), radius, mat)

    @classmethod
    def moving(cls, center1, center2, radius, mat):
        return cls(_Path(center1, center2 - center1), radius, mat)


class quad:
    def __init__(self, Q, u, v, mat):
        self.Q, self.u, self.v, self.mat = Q, u, v, mat
        n = u.cross(v)
        self.normal = n.unit_vector()
        self.D = self.normal.dot(Q)
        self.w = n / n.dot(n)


class triangle:
    def __init__(self, v0, v1, v2, mat):
        self.v0, self.v1, self.v2, self.mat = v0, v1, v2, mat
        self.edge1 = v1 - v0
        self.edge2 = v2 - v0
        self.normal = self.edge1.cross(self.edge2).unit_vector()


class mesh:
    """Triangle mesh from a Wavefront OBJ file, with the reference's constructor
    (core/mesh.py:25-62: model folder or file, material, scale, offset) but its own parser -- no
    pywavefront -- and array storage: vertices are transformed as `v * scale + offset`
    (core/mesh.py:207-222), polygons fan-triangulated (:134-139), zero-area triangles dropped
    (:224-246).  A million-triangle mesh stays three numpy arrays; `triangles` materialises
    `triangle` objects only if somebody asks for them."""

    def __init__(self, model_path, mat, scale: float = 1.0, offset=None, obj_filename=None, use_bvh: bool = True):
        import os
        self.mat, self.scale, self.offset = mat, scale, offset if offset is not None else point3(0, 0, 0)
        path = str(model_path)
        if os.path.isdir(path):
            found = None
            for root, _, files in sorted(os.walk(path)):
                for f in sorted(files):
                    if f.lower().endswith(".obj") and (obj_filename is None or f == obj_filename):
                        found = os.path.join(root, f)
                        break
                if found:
                    break
            if not found:
                raise FileNotFoundError(f"No OBJ files found in {path}")
            path = found
        elif not os.path.exists(path):
            raise FileNotFoundError(f"Model path does not exist: {path}")
        self.obj_file = path
        self._load(path)

    @classmethod
    def from_arrays(cls, v0, v1, v2, mat):
        self = cls.__new__(cls)
        self.mat, self.scale, self.offset, self.obj_file = mat, 1.0, point3(0, 0, 0), None
        self._set(np.asarray(v0, np.float64), np.asarray(v1, np.float64), np.asarray(v2, np.float64))
        return self

    def _load(self, path):
        verts, faces = [], []
        with open(path) as f:
            for line in f:
                if line.startswith("v "):
                    p = line.split()
                    verts.append((float(p[1]), float(p[2]), float(p[3])))
                elif line.startswith("f "):
                    idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                    idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                    for k in range(1, len(idx) - 1):
                        faces.append((idx[0], idx[k], idx[k + 1]))
        if not faces:
            raise ValueError("No valid triangles created from OBJ file")
        V = np.asarray(verts, np.float64) * self.scale + np.array([self.offset.x, self.offset.y, self.offset.z])
        F = np.asarray(faces, np.int64)
        self._set(V[F[:, 0]], V[F[:, 1]], V[F[:, 2]])

    def _set(self, v0, v1, v2):
        cr = np.cross(v1 - v0, v2 - v0)
        keep = (cr ** 2).sum(axis=1) >= 1e-10
        self.v0, self.v1, self.v2 = v0[keep], v1[keep], v2[keep]
        self.num_degenerate = int((~keep).sum())
        if not len(self.v0):
            raise ValueError("No valid triangles created from OBJ file")
        self._triangles = None

    def tri_arrays(self):
        """(v0, v1, v2) float64 arrays -- the scene compiler's fast path."""
        return self.v0, self.v1, self.v2

    @property
    def triangles(self):
        if self._triangles is None:
            self._triangles = [triangle(point3(*a), point3(*b), point3(*c), self.mat)
                               for a, b, c in zip(self.v0, self.v1, self.v2)]
        return self._triangles


class hittable_list:
    def __init__(self, objects=None):
        self.objects = list(objects) if objects else []

    def add(self, obj):
        self.objects.append(obj)


class bvh_node:
    """Container with the reference's constructor shape (core/bvh_node.py:13-46: `from_objects(objects, start,
    end)`).  The reference sorts the span along its longest axis and splits it in the middle to make a
    traversable object tree; the Taichi path never traverses that tree -- its compiler only walks it to collect
    the primitives (scene_compiler.py:28-146) and then builds its own SAH tree -- so here the span is split in
    the middle as it stands.  Same primitives, same SAH input set; only their order (hence their indices) differs."""

    def __init__(self, left=None, right=None):
        self.left, self.right = left, right

    @classmethod
    def from_objects(cls, objects, start: int, end: int):
        span = end - start
        if span <= 0:
            return cls(None, None)
        if span == 1:
            return cls(objects[start], None)
        if span == 2:
            return cls(objects[start], objects[start + 1])
        mid = start + span // 2
        return cls(cls.from_objects(objects, start, mid), cls.from_objects(objects, mid, end))


class constant_medium:
    def __init__(self, boundary, phase_function, density):
        self.boundary = boundary
        self.phase_function = phase_function
        self.neg_inv_density = -1 / density

    @classmethod
    def from_color(cls, boundary, albedo, density):
        return cls(boundary, isotropic.from_color(albedo), density)


def box(a, b, mat, angle: float = 0.0) -> hittable_list:
    """Six quads of an axis-aligned box, optionally rotated about Y through its centre by `angle`
    degrees (same faces, order and winding as the reference's scenes.box, src/scenes.py:961-1024)."""
    lo = vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z))
    hi = vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z))
    dx, dy, dz = vec3(hi.x - lo.x, 0, 0), vec3(0, hi.y - lo.y, 0), vec3(0, 0, hi.z - lo.z)
    corners = [(vec3(lo.x, lo.y, hi.z), dx, dy), (vec3(hi.x, lo.y, hi.z), -dz, dy), (vec3(hi.x, lo.y, lo.z), -dx, dy),
               (vec3(lo.x, lo.y, lo.z), dz, dy), (vec3(lo.x, hi.y, hi.z), dx, -dz), (vec3(lo.x, lo.y, lo.z), dx, dz)]
    if angle != 0.0:
        th = math.radians(angle)
        c, s = math.cos(th), math.sin(th)
        rot = lambda p: vec3(c * p.x + s * p.z, p.y, -s * p.x + c * p.z)
        ctr = (lo + hi) * 0.5
        corners = [(rot(q - ctr) + ctr, rot(u), rot(v)) for q, u, v in corners]
    return hittable_list([quad(q, u, v, mat) for q, u, v in corners])


class camera:
    aspect_ratio = 1.0
    img_width = 100
    samples_per_pixel = 10
    vfov = 90
    lookfrom = point3(0, 0, 0)
    lookat = point3(0, 0, -1)
    vup = vec3(0, 1, 0)
    defocus_angle = 0.0
    focus_distance = 10.0

    def initialize(self):
        from .scenes import camera_vectors
        cam, h = camera_vectors([self.lookfrom.x, self.lookfrom.y, self.lookfrom.z],
                                [self.lookat.x, self.lookat.y, self.lookat.z], [self.vup.x, self.vup.y, self.vup.z],
                                self.vfov, self.img_width, self.aspect_ratio, self.defocus_angle, self.focus_distance)
        self.img_height = h
        self.center = vec3(*cam["cam_center"])
        self.pixel00_loc = vec3(*cam["cam_pixel00"])
        self.delta_u, self.delta_v = vec3(*cam["cam_delta_u"]), vec3(*cam["cam_delta_v"])
        self.defocus_disk_u, self.defocus_disk_v = vec3(*cam["cam_defocus_disk_u"]), vec3(*cam["cam_defocus_disk_v"])
