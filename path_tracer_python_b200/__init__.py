"""path_tracer_python_b200 -- the Taichi path-tracing hot path of fakhirsh/path-tracer-python,
rebuilt for B200 (sm_100a) behind the reference's renderer API.

Importing the package is cheap and GPU-free; the CUDA library is loaded on first use and there is
no CPU fallback (see _lib.lib()).
"""
__all__ = ["TaichiRenderer", "B200Renderer", "RendererFactory"]


def __getattr__(name):
    if name in ("TaichiRenderer", "B200Renderer"):
        from .renderer import TaichiRenderer
        return TaichiRenderer
    if name == "RendererFactory":
        from .renderer_factory import RendererFactory
        return RendererFactory
    raise AttributeError(name)
