"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

Imports the *host side* of the reference (fakhirsh/path-tracer-python, mounted read-only at
/root/reference) inside this container so that golden fixtures can be generated from the
reference's own code:

  * scene object graphs from ``src/scenes.py`` (captured at the point where the scene function
    hands ``(world, cam)`` to ``RendererFactory.create`` / ``InteractiveViewer``),
  * ``compile_scene`` (src/render_server/taichi_renderer/scene_compiler.py:931) and
    ``compile_bvh`` (src/render_server/taichi_renderer/bvh_compiler.py:132) outputs,
  * the Perlin tables the renderer builds right after (renderer.py:78-80, core/perlin.py:7-55).

``taichi``, ``tkinter``, ``pywavefront`` and ``PIL.ImageTk`` are absent from this image; they are
stubbed (the host-side code never calls into them).  For executing the reference's *device* code
(kernels.py) see ``oracle/taichi_shim``.

Nothing here may be imported by the product package, ``bench.py`` or the ``-m gpu`` tests:
/root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import contextlib
import io
import os
import random
import sys
import types
from unittest import mock

import numpy as np

REF_ROOT = os.environ.get("PT_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "render_server", "taichi_renderer"))


class SceneCaptured(Exception):
    def __init__(self, world, cam, img_path):
        super().__init__("scene captured")
        self.world, self.cam, self.img_path = world, cam, img_path


_imported = False


def import_reference_host(stub_taichi: bool = True):
    """Put the reference's ``src`` on sys.path with GUI/JIT dependencies stubbed."""
    global _imported
    if _imported:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    if stub_taichi and "taichi" not in sys.modules:
        ti = mock.MagicMock(name="taichi")
        sys.modules["taichi"] = ti
        sys.modules["taichi.math"] = ti.math
    for name in ("tkinter", "pywavefront"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mock.MagicMock(name=name)
    try:
        import PIL.ImageTk  # noqa: F401
    except Exception:
        sys.modules["PIL.ImageTk"] = mock.MagicMock(name="PIL.ImageTk")
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    _imported = True


def capture_scene(scene_name: str, seed: int = 1234):
    """Run ``scenes.<scene_name>()`` until it constructs its renderer; return (world, cam).

    The reference scenes use the global ``random`` stream (src/scenes.py:1166, :1217, :1448), so
    it is seeded right before the call, as SURVEY.md section 8(d) prescribes.
    """
    import_reference_host()
    cwd = os.getcwd()
    os.chdir(REF_SRC)  # image_texture("assets/images/earthmap.jpg") is cwd-relative
    try:
        import scenes  # type: ignore
        import render_server.interactive_viewer as iv  # type: ignore

        def _capture(*args, **kwargs):
            # RendererFactory.create(type, world, cam, path) / InteractiveViewer(world, cam, path)
            if isinstance(args[0], str):
                args = args[1:]
            raise SceneCaptured(args[0], args[1], args[2] if len(args) > 2 else None)

        fake_factory = types.SimpleNamespace(create=_capture)
        with mock.patch.object(scenes, "RendererFactory", fake_factory), \
                mock.patch.object(scenes, "InteractiveViewer", _capture), \
                mock.patch.object(iv, "InteractiveViewer", _capture):
            random.seed(seed)
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    getattr(scenes, scene_name)()
            except SceneCaptured as cap:
                return cap.world, cap.cam
        raise RuntimeError(f"scene {scene_name} never constructed a renderer")
    finally:
        os.chdir(cwd)


def reference_compile(world):
    """Run the reference's own compile_scene + compile_bvh; returns the canonical flat dict used
    everywhere in this repo (see path_tracer_python_b200/scene_compiler.py: CompiledScene)."""
    import_reference_host()
    from render_server.taichi_renderer.scene_compiler import compile_scene  # type: ignore
    from render_server.taichi_renderer.bvh_compiler import compile_bvh  # type: ignore

    with contextlib.redirect_stdout(io.StringIO()):
        (geo, mat, spheres, qgeo, qmat, quads, tgeo, tmat, tris, _reg, img_list) = compile_scene(world)
        bvh = compile_bvh(world, spheres, quads, tris)
    out = {}
    out["sphere_data"] = geo["sphere_data"]
    for k in ("quad_Q", "quad_u", "quad_v", "quad_normal", "quad_D", "quad_w"):
        out[k] = qgeo[k]
    for k in ("triangle_v0", "triangle_v1", "triangle_v2", "triangle_edge1", "triangle_edge2",
              "triangle_normal"):
        out[k] = tgeo[k]
    for prefix, d in (("sphere", mat), ("quad", qmat), ("triangle", tmat)):
        for k, v in d.items():
            out[f"{prefix}_{k}"] = v
    for k in ("bvh_bbox_min", "bvh_bbox_max", "bvh_left_child", "bvh_right_child", "bvh_parent",
              "bvh_prim_type", "bvh_prim_idx"):
        out[k] = np.asarray(bvh[k])
    out["num_images"] = np.int32(len(img_list))
    for i, tex in enumerate(img_list):
        f = tex.image.fdata
        u8 = np.rint(f * 255.0).astype(np.uint8)
        # the device stores 8-bit texels and divides by 255.0f on read; make sure that
        # reproduces the reference's float32 ``fdata`` (util/rtw_image.py:66) bit for bit
        assert np.array_equal(u8.astype(np.float32) / np.float32(255.0), f)
        out[f"image_{i}"] = u8
    return out


def reference_perlin_tables():
    """A fresh ``perlin()`` exactly as TaichiRenderer.__init__ makes one (renderer.py:78-80)."""
    import_reference_host()
    from core.perlin import perlin  # type: ignore

    p = perlin()
    rv = np.array([[v.x, v.y, v.z] for v in p.randvec], dtype=np.float32)
    return {
        "perlin_randvec": rv,
        "perlin_perm_x": np.array(p.perm_x, dtype=np.int32),
        "perlin_perm_y": np.array(p.perm_y, dtype=np.int32),
        "perlin_perm_z": np.array(p.perm_z, dtype=np.int32),
    }


def camera_arrays(cam):
    """cam.initialize() (core/camera.py:34-72) then the vectors _upload_camera uploads
    (renderer.py:230-247), kept in float64 (the narrowing to f32 happens at upload)."""
    cam.initialize()
    v = lambda a: [a.x, a.y, a.z]
    return {
        "cam_center": np.array(v(cam.center), dtype=np.float64),
        "cam_pixel00": np.array(v(cam.pixel00_loc), dtype=np.float64),
        "cam_delta_u": np.array(v(cam.delta_u), dtype=np.float64),
        "cam_delta_v": np.array(v(cam.delta_v), dtype=np.float64),
        "cam_defocus_disk_u": np.array(v(cam.defocus_disk_u), dtype=np.float64),
        "cam_defocus_disk_v": np.array(v(cam.defocus_disk_v), dtype=np.float64),
        "cam_defocus_angle": np.float64(cam.defocus_angle),
        "img_width": np.int32(cam.img_width),
        "img_height": np.int32(cam.img_height),
    }
