"""Output step: accumulation buffer -> 8-bit image (reference preview.py:117-132), and an optional
Tk live preview with the reference's LivePreview interface (tkinter is imported lazily; it is not
installed on the GPU boxes)."""
from __future__ import annotations

import numpy as np


def buffer_to_image(accum_buffer: np.ndarray, sample_count: int) -> np.ndarray:
    """u8 = clip(sqrt(max(0, accum / spp)) * 255.999, 0, 255).  Host/numpy version, kept for API
    parity; the renderer itself tonemaps on the device (pt_tonemap)."""
    scale = np.float32(1.0 / max(1, sample_count))
    scaled = np.asarray(accum_buffer, np.float32) * scale
    g = np.sqrt(np.maximum(np.float32(0), scaled))
    return np.clip(g * np.float32(255.999), 0, 255).astype(np.uint8)


class LivePreview:
    def __init__(self, width: int, height: int, title: str = "Path Tracer - Live Preview"):
        self.width, self.height, self.title = width, height, title
        self.window = None

    def start(self, update_interval_ms: int = 50):
        import tkinter as tk  # noqa: F401  (raises ImportError where there is no display stack)
        self._tk = tk
        self.window = tk.Tk()
        self.window.title(self.title)
        self.label = tk.Label(self.window)
        self.label.pack()
        self._photo = None

    def update(self, accum_np: np.ndarray, current_sample: int, total_samples: int | None = None):
        if self.window is None:
            return
        from PIL import Image, ImageTk
        img = Image.fromarray(buffer_to_image(accum_np, current_sample), mode="RGB")
        self._photo = ImageTk.PhotoImage(img)
        self.label.config(image=self._photo)

    def update_image(self, img_u8: np.ndarray):
        """Show an already tonemapped 8-bit frame (the renderer tonemaps on the device)."""
        if self.window is None:
            return
        from PIL import Image, ImageTk
        self._photo = ImageTk.PhotoImage(Image.fromarray(img_u8, mode="RGB"))
        self.label.config(image=self._photo)

    def process_events(self):
        if self.window is not None:
            try:
                self.window.update()
            except Exception:
                self.window = None

    def finish(self):
        if self.window is not None:
            try:
                self.window.mainloop()
            except Exception:
                pass

    def close(self):
        if self.window is not None:
            try:
                self.window.destroy()
            except Exception:
                pass
            self.window = None

    buffer_to_image = staticmethod(buffer_to_image)
