"""RendererFactory with the reference's interface (src/render_server/renderer_factory.py:20-44):
``RendererFactory.create(renderer_type, world, cam, img_path, **kwargs)``.  'taichi' (the name the
reference's scenes use, e.g. src/scenes.py:1504) and 'b200' both resolve to the CUDA renderer."""
from __future__ import annotations

from .renderer import TaichiRenderer


class RendererFactory:
    _renderers = {"taichi": TaichiRenderer, "b200": TaichiRenderer}

    @classmethod
    def create(cls, renderer_type: str, world, cam, img_path: str, **kwargs):
        renderer_type = renderer_type.lower()
        if renderer_type not in cls._renderers:
            available = ", ".join(cls._renderers.keys())
            raise ValueError(f"Unknown renderer type '{renderer_type}'. Available: {available}")
        return cls._renderers[renderer_type](world, cam, img_path, **kwargs)

    @classmethod
    def get_available_renderers(cls) -> list:
        return list(cls._renderers.keys())
