"""Engine: the thin layer between the reference-shaped renderer class and the C ABI.

PyTorch is used only for memory (pinned host blobs, device buffers, streams); every kernel is
libb200pt's.  One Engine = one PtContext = one GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import HostScene, PtCamera, PtCounters, PtPackedInfo, check, make_camera


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.B200PTError("no CUDA device: path_tracer_python_b200 has no CPU path (by design)")


class Engine:
    def __init__(self, device: int | None = None, numerics: str = "strict"):
        """numerics: "strict" = the PT-F32 contract (bit-exact against the reference's kernels; the parity
        build) or "fast" = the production build of the same sources (FMA contraction and approximate reciprocal /
        division / square root around the strict intersection tests; first hits equal to strict's)."""
        _require_cuda()
        self.numerics = numerics
        self.L = _lib.lib(numerics)
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        ctx = C.c_void_p()
        self._check(self.L.pt_create(self.device_index, C.byref(ctx)), "pt_create")
        self.ctx = ctx
        self.info: PtPackedInfo | None = None
        self.scene_blob = None          # device uint8 tensor (owned here, borrowed by the library)
        self._pinned_blob = None
        self._upload_done = None         # event: the last pinned -> device scene copy has finished
        self._workspace = None
        self.h2d_bytes = 0
        self.wavefront_samples_in_flight = None  # None = auto (about 66 M paths per batch); bounded by max_workspace_bytes
        self.max_workspace_bytes = 24 << 30        # two pipelines x ~66 M paths x 168 B
        self.wavefront_pipelines = 2             # batches in flight at once (option "wf_pipelines"): each has its own queue set
        self.wavefront_min_batches = 1           # auto sizing: split a render call into at least this many batches

    def close(self):
        if getattr(self, "ctx", None):
            self.L.pt_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str = ""):
        check(rc, what, self.L)

    @staticmethod
    def has_numerics(numerics: str) -> bool:
        return _lib.has_numerics(numerics)

    # ------------------------------------------------------------------ configuration
    def set_option(self, name: str, value: int):
        if name == "wf_pipelines":
            self.wavefront_pipelines = int(value)
        self._check(self.L.pt_set_option(self.ctx, name.encode(), int(value)), f"pt_set_option({name})")

    def set_row_interleave(self, block_rows: int = 0, stride_rows: int = 0):
        """Interleaved row blocks for the multi-GPU tile split (pt_set_row_interleave): afterwards
        render_*(…, y0=rank * block_rows, y1=y0 + local_rows) renders this rank's blocks in one call.
        block_rows = 0 switches back to contiguous windows."""
        self._check(self.L.pt_set_row_interleave(self.ctx, int(block_rows), int(stride_rows)), "pt_set_row_interleave")

    def _stream_ptr(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ scene / camera
    def set_scene(self, cs: dict):
        """cs: CompiledScene dict (reference layout).  Packs on the host (C++), one pinned->device copy."""
        hs = HostScene(cs)
        info = PtPackedInfo()
        self._check(self.L.pt_pack_scene_info(C.byref(hs.c), C.byref(info)), "pt_pack_scene_info")
        if int(info.max_depth_tree) > 64:      # what pt_set_scene would refuse: do not upload, keep the current scene
            raise _lib.B200PTError(f"BVH depth {int(info.max_depth_tree)} exceeds the 64-entry traversal stack (kernels.py:649)")
        nbytes = int(info.total_bytes)
        # the pinned staging buffer is reused: the previous upload must have left it before it is rewritten
        if self._upload_done is not None:
            self._upload_done.synchronize()
        if self._pinned_blob is None or self._pinned_blob.numel() < nbytes:
            self._pinned_blob = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        pinned = self._pinned_blob[:nbytes]
        self._check(self.L.pt_pack_scene(C.byref(hs.c), C.byref(info), C.c_void_p(pinned.data_ptr())), "pt_pack_scene")
        with torch.cuda.device(self.device):
            blob = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            blob.copy_(pinned, non_blocking=True)
            self._upload_done = torch.cuda.Event()
            self._upload_done.record()
        self._check(self.L.pt_set_scene(self.ctx, C.c_void_p(blob.data_ptr()), C.byref(info)), "pt_set_scene")
        self.scene_blob, self.info = blob, info
        self.h2d_bytes = nbytes
        return info

    def reupload_scene(self):
        """Copy the already packed scene from pinned host memory to the device again (the per-frame
        host -> device input copy of `render_frame`; packing happened once in set_scene)."""
        if self.info is None or self._pinned_blob is None:
            raise _lib.B200PTError("reupload_scene: no scene has been set")
        nbytes = int(self.info.total_bytes)
        with torch.cuda.device(self.device):
            self.scene_blob.copy_(self._pinned_blob[:nbytes], non_blocking=True)
            self._upload_done = torch.cuda.Event()
            self._upload_done.record()
        return nbytes

    def set_camera(self, cam: dict | PtCamera, background=(0.7, 0.8, 1.0), max_depth=50, jitter: str = "random"):
        """jitter: "random" = the sub-pixel offset is two draws of the path's random stream (the reference);
        "sobol" = low-discrepancy offsets, the sample's point of a per-pixel scrambled (0,2)-sequence."""
        pc = cam if isinstance(cam, PtCamera) else make_camera(cam, background, max_depth, jitter)
        self._check(self.L.pt_set_camera(self.ctx, C.byref(pc)), "pt_set_camera")
        self.camera = pc

    # ------------------------------------------------------------------ rendering
    def new_accum(self, W: int, H: int) -> torch.Tensor:
        return torch.zeros((H, W, 3), dtype=torch.float32, device=self.device)

    def render_megakernel(self, accum: torch.Tensor, sample_begin: int, sample_count: int, seed: int = 0,
                          y0: int = 0, y1: int | None = None):
        H, W, _ = accum.shape
        assert accum.is_cuda and accum.dtype == torch.float32 and accum.is_contiguous()
        y1 = H if y1 is None else y1
        with torch.cuda.device(self.device):
            self._check(self.L.pt_render_megakernel(self.ctx, C.c_void_p(accum.data_ptr()), W, H, y0, y1, sample_begin,
                                              sample_count, C.c_uint32(seed & 0xFFFFFFFF), self._stream_ptr()),
                  "pt_render_megakernel")

    def render_wavefront(self, accum: torch.Tensor, sample_begin: int, sample_count: int, seed: int = 0,
                         y0: int = 0, y1: int | None = None):
        H, W, _ = accum.shape
        assert accum.is_cuda and accum.dtype == torch.float32 and accum.is_contiguous()
        y1 = H if y1 is None else y1
        if self.wavefront_samples_in_flight is None:
            # per-batch samples, times the two pipelines the library runs consecutive batches on (two queue sets)
            b = auto_samples_in_flight(W, y1 - y0, sample_count, min_batches=self.wavefront_min_batches)
            if b < sample_count:
                b *= self.wavefront_pipelines
        else:
            b = max(1, min(self.wavefront_samples_in_flight, sample_count))
        # room for the node copy a call builds in the workspace for its primary wave: origin-relative boxes in eight
        # direction-octant variants (pt_render_wavefront; one copy when the tree is large)
        nb = int(self.info.node_bytes) * int(self.info.num_inner) if self.info is not None else 0
        extra = (8 * nb + 2048) if 8 * nb + 2048 <= (1 << 30) else nb + 256
        need = int(self.L.pt_wavefront_workspace_bytes_for(W, y1 - y0, b)) + extra
        while b > 1 and need > self.max_workspace_bytes:
            b -= 1
            need = int(self.L.pt_wavefront_workspace_bytes_for(W, y1 - y0, b)) + extra
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.L.pt_render_wavefront(self.ctx, C.c_void_p(accum.data_ptr()), W, H, y0, y1, sample_begin,
                                             sample_count, C.c_uint32(seed & 0xFFFFFFFF),
                                             C.c_void_p(self._workspace.data_ptr()), C.c_size_t(need),
                                             self._stream_ptr()), "pt_render_wavefront")

    # ------------------------------------------------------------------ wavefront, one stage at a time
    def wf_begin(self, W: int, H: int, sample_base: int, samples: int, seed: int = 0, y0: int = 0, y1: int | None = None):
        """Bind one batch (`samples` samples of every pixel of rows [y0, y1)) for the stage-level calls below --
        the reference's wavefront kernels one at a time, in the order its host loop launches them
        (renderer.py:305-334): wf_generate, then per bounce wf_extend + wf_shade until wf_compact says no ray
        is left, then wf_resolve."""
        y1 = H if y1 is None else y1
        need = int(self.L.pt_wavefront_workspace_bytes_for(W, y1 - y0, samples))
        if getattr(self, "_stage_ws", None) is None or self._stage_ws.numel() < need:
            self._stage_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        self._stage_paths = W * (y1 - y0) * samples
        self._check(self.L.pt_wf_begin(self.ctx, W, H, y0, y1, int(sample_base), int(samples), C.c_uint32(seed & 0xFFFFFFFF),
                                       C.c_void_p(self._stage_ws.data_ptr()), C.c_size_t(need)), "pt_wf_begin")

    def wf_generate(self):
        with torch.cuda.device(self.device):
            self._check(self.L.pt_wf_generate(self.ctx, self._stream_ptr()), "pt_wf_generate")

    def wf_extend(self, bounce: int):
        with torch.cuda.device(self.device):
            self._check(self.L.pt_wf_extend(self.ctx, int(bounce), self._stream_ptr()), "pt_wf_extend")

    def wf_shade(self, bounce: int):
        with torch.cuda.device(self.device):
            self._check(self.L.pt_wf_shade(self.ctx, int(bounce), self._stream_ptr()), "pt_wf_shade")

    def wf_compact(self, bounce: int) -> int:
        """Live rays in the queue wave `bounce` reads (synchronises; the reference's active_ray_count read-back)."""
        n = C.c_int(0)
        with torch.cuda.device(self.device):
            self._check(self.L.pt_wf_compact(self.ctx, int(bounce), C.byref(n), self._stream_ptr()), "pt_wf_compact")
        return int(n.value)

    def wf_read_rays(self, bounce: int, n: int) -> dict:
        f = lambda: torch.empty((n, 3), dtype=torch.float32, device=self.device)
        out = {"o": f(), "d": f(), "throughput": f(), "pixel": torch.empty(n, dtype=torch.int32, device=self.device),
               "depth": torch.empty(n, dtype=torch.int32, device=self.device)}
        with torch.cuda.device(self.device):
            self._check(self.L.pt_wf_read_rays(self.ctx, int(bounce), int(n), *[C.c_void_p(out[k].data_ptr()) for k in
                                               ("o", "d", "throughput", "pixel", "depth")], self._stream_ptr()), "pt_wf_read_rays")
        return out

    def wf_read_hits(self, n: int) -> dict:
        out = {"t": torch.empty(n, dtype=torch.float32, device=self.device), "prim": torch.empty(n, dtype=torch.int32, device=self.device),
               "t_exit": torch.empty(n, dtype=torch.float32, device=self.device)}
        with torch.cuda.device(self.device):
            self._check(self.L.pt_wf_read_hits(self.ctx, int(n), *[C.c_void_p(out[k].data_ptr()) for k in ("t", "prim", "t_exit")],
                                               self._stream_ptr()), "pt_wf_read_hits")
        return out

    def wf_resolve(self, accum: torch.Tensor):
        assert accum.is_cuda and accum.dtype == torch.float32 and accum.is_contiguous()
        with torch.cuda.device(self.device):
            self._check(self.L.pt_wf_resolve(self.ctx, C.c_void_p(accum.data_ptr()), self._stream_ptr()), "pt_wf_resolve")

    # ------------------------------------------------------------------ probes
    def generate_camera_rays(self, W: int, H: int, seed: int = 0, sample: int = 0):
        o = torch.empty((H * W, 3), dtype=torch.float32, device=self.device)
        d = torch.empty_like(o)
        with torch.cuda.device(self.device):
            self._check(self.L.pt_generate_camera_rays(self.ctx, W, H, C.c_uint32(seed & 0xFFFFFFFF), sample,
                                                 C.c_void_p(o.data_ptr()), C.c_void_p(d.data_ptr()), self._stream_ptr()),
                  "pt_generate_camera_rays")
        return o, d

    def first_hit(self, o: torch.Tensor, d: torch.Tensor, t_min=0.001, t_max=1e10, want_pn=True):
        o = o.to(self.device, torch.float32).contiguous().view(-1, 3)
        d = d.to(self.device, torch.float32).contiguous().view(-1, 3)
        n = o.shape[0]
        t = torch.empty(n, dtype=torch.float32, device=self.device)
        prim = torch.empty(n, dtype=torch.int32, device=self.device)
        p = torch.empty((n, 3), dtype=torch.float32, device=self.device) if want_pn else None
        nn = torch.empty((n, 3), dtype=torch.float32, device=self.device) if want_pn else None
        with torch.cuda.device(self.device):
            self._check(self.L.pt_first_hit(self.ctx, C.c_void_p(o.data_ptr()), C.c_void_p(d.data_ptr()), n, t_min, t_max,
                                      C.c_void_p(t.data_ptr()), C.c_void_p(prim.data_ptr()),
                                      C.c_void_p(p.data_ptr()) if want_pn else None,
                                      C.c_void_p(nn.data_ptr()) if want_pn else None, self._stream_ptr()), "pt_first_hit")
        return t, prim, p, nn

    def math_eval(self, fn: int, a: torch.Tensor, b: torch.Tensor | None = None):
        a = a.to(self.device, torch.float32).contiguous()
        b = a if b is None else b.to(self.device, torch.float32).contiguous()
        out = torch.empty_like(a)
        with torch.cuda.device(self.device):
            self._check(self.L.pt_math_eval(self.ctx, fn, a.numel(), C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()),
                                      C.c_void_p(out.data_ptr()), self._stream_ptr()), "pt_math_eval")
        return out

    def probe_gather(self, working_set_bytes: int, pattern: int = 0, iters: int = 64, repeats: int = 5) -> float:
        """Measured GB/s of 64-byte record gathers (the node-fetch instruction sequence, all lanes
        active) over a zero-filled working set of that size: the memory roof of the traversal kernels
        for a scene of that size.  pattern 0 = random record per lane, 1 = one record per warp."""
        buf = torch.zeros(max(64, int(working_set_bytes)) // 64 * 64, dtype=torch.uint8, device=self.device)
        out = C.c_double(0.0)
        with torch.cuda.device(self.device):
            self._check(self.L.pt_probe_gather(self.ctx, C.c_void_p(buf.data_ptr()), C.c_size_t(buf.numel()), int(pattern),
                                         int(iters), int(repeats), C.byref(out), self._stream_ptr()), "pt_probe_gather")
        return float(out.value)

    def tonemap(self, accum: torch.Tensor, sample_count: int) -> torch.Tensor:
        out = torch.empty(accum.shape, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.L.pt_tonemap(self.ctx, C.c_void_p(accum.data_ptr()), accum.numel(), int(sample_count),
                                    C.c_void_p(out.data_ptr()), self._stream_ptr()), "pt_tonemap")
        return out

    # ------------------------------------------------------------------ statistics
    def debug_error(self) -> int:
        """First bounds violation the kernels of the debug build recorded since the last call
        (check code << 24 | source line; 0 = none).  Always 0 with the product builds."""
        code = C.c_uint32(0)
        self._check(self.L.pt_get_debug_error(self.ctx, C.byref(code)), "pt_get_debug_error")
        return int(code.value)

    def enable_counters(self, on: bool = True):
        self._check(self.L.pt_enable_counters(self.ctx, 1 if on else 0), "pt_enable_counters")

    def counters(self, reset: bool = True) -> dict:
        c = PtCounters()
        self._check(self.L.pt_get_counters(self.ctx, C.byref(c), 1 if reset else 0), "pt_get_counters")
        return c.as_dict()


    def enable_stage_timing(self, on: bool = True):
        self._check(self.L.pt_enable_stage_timing(self.ctx, 1 if on else 0), "pt_enable_stage_timing")

    def stage_times(self) -> list:
        """[(stage name, wave, ms)] for every launch of the last render_wavefront call (synchronises)."""
        from ._lib import PtStageTime, STAGE_NAMES
        n = C.c_int(0)
        self._check(self.L.pt_get_stage_times(self.ctx, None, 0, C.byref(n)), "pt_get_stage_times")
        buf = (PtStageTime * max(1, n.value))()
        self._check(self.L.pt_get_stage_times(self.ctx, buf, n.value, C.byref(n)), "pt_get_stage_times")
        return [(STAGE_NAMES[buf[i].kernel], buf[i].wave, buf[i].ms) for i in range(n.value)]


    def stage_counters(self) -> list:
        """Per-launch counter increments of the last render_wavefront call (needs enable_counters and
        enable_stage_timing); same order as stage_times()."""
        n = C.c_int(0)
        self._check(self.L.pt_get_stage_counters(self.ctx, None, 0, C.byref(n)), "pt_get_stage_counters")
        buf = (PtCounters * max(1, n.value))()
        self._check(self.L.pt_get_stage_counters(self.ctx, buf, n.value, C.byref(n)), "pt_get_stage_counters")
        cum = [buf[i].as_dict() for i in range(n.value)]
        out, prev = [], None
        for c in cum:
            out.append({k: c[k] - (prev[k] if prev else 0) for k in c})
            prev = c
        return out


def auto_samples_in_flight(W: int, rows: int, sample_count: int, target_paths: float = 66.4e6, min_batches: int = 1) -> int:
    """Samples of every pixel the wavefront pipeline keeps in flight per batch: about `target_paths`
    paths (32 samples of a 1080p frame -- fewer leave the later waves too small to fill the GPU, more
    gain nothing), split into equal batches, at most 256 (the library's limit).  A call that would fit
    ONE batch is still split in two when both halves keep at least 30 M paths (the share of one of 8
    GPUs at 1080p x 256 spp): the first batch's tail -- a latency chain that leaves the machine almost
    empty -- then runs under the second batch's primary wave (measured +2.2 % at 32 samples; four
    batches of 8 are 5 % slower again).  `min_batches` forces a finer split."""
    n_pix = max(1, W * rows)
    per_batch = target_paths / n_pix
    n_batches = max(1, round(sample_count / per_batch))
    if n_batches == 1 and sample_count >= 2 and (sample_count // 2) * n_pix >= 30e6:
        n_batches = 2
    n_batches = max(n_batches, min(int(min_batches), max(1, sample_count)))
    return max(1, min(256, -(-sample_count // n_batches)))


def algorithmic_bytes(counters: dict) -> int:
    """32 B per child-box test + primitive record bytes per leaf test."""
    return (32 * counters["box_tests"] + 16 * counters["sphere_tests"] + 48 * counters["triangle_tests"]
            + 64 * counters["quad_tests"])
