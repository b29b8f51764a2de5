#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 path-tracing hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode megakernel|wavefront]
                    [--scaling auto|strong|weak] [--numerics strict|fast]

Workload (BASELINE.json configs[1]): vol2_final_scene (1007 spheres, 2401 quads, 6815-node SAH BVH,
fog dome, Perlin + image textures), 1920x1080, 256 spp, max_depth 50, black background.
A *step* = one full render of that frame: W*H*256 paths through the wavefront pipeline (or the megakernel).

  value  : Mpaths/s, whole job, scene blob + accumulation buffer resident in HBM, CUDA events.
  e2e    : Mpaths/s through the reference-facing API TaichiRenderer.render_frame(): packed scene
           copied pinned-host -> device, camera upload, render, (N > 1: NCCL reduce,) accumulation
           buffer copied back to pinned host memory, all inside the timed region.
  N > 1  : one process per GPU (torchrun).  Default = STRONG scaling (north_star: "scales >= 7x from 1
           to 8 GPUs" on the fixed frame): the 256 samples are split across the ranks (rank r renders
           samples [r*256/N, (r+1)*256/N) of every pixel; streams are keyed by (seed, pixel, sample), so
           the union is exactly the 1-GPU sample set) and the fp32 accumulation buffers are summed onto
           rank 0 with one NCCL reduce inside the timed region.  Render and reduce times are reported
           per rank; rank 0 also renders the frame alone and reports the difference to the reduced
           buffer.  --scaling weak keeps 256 samples per rank (per-GPU work fixed).
  config5: BASELINE.json configs[4] (3840x2160, 1024 spp, same sharding), one timed step, in the same line.
  --impl reference : the CPU oracle (a scalar C++/OpenMP restatement of the reference's Taichi
           kernels -- Taichi itself is not installable here) on ALL host cores (the thread count is set
           explicitly: torchrun exports OMP_NUM_THREADS=1), on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for p in (REPO, os.path.join(REPO, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "Mpaths/s vol2_final_scene 1920x1080 256spp"
WIDTH, ASPECT, SPP, MAX_DEPTH, BG = 1920, 16.0 / 9.0, 256, 50, (0.0, 0.0, 0.0)
SEED = 1234
C5_WIDTH, C5_SPP = 3840, 1024


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def host_cores() -> int:
    """Cores this process may run on (what the CPU arm uses, whatever OMP_NUM_THREADS says)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks + throttle reasons, sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(cs, target_seconds=12.0, mode="megakernel"):
    """Time the CPU oracle on a bounded sample of the workload: the same scene and camera at
    480x270, 8 spp per call, repeated (fresh sample indices each call) until ~target_seconds of
    CPU work have been measured.  Returns (Mpaths/s, cores, sample description)."""
    import oracle as O
    from path_tracer_python_b200 import scenes
    O.build()
    nthreads = host_cores()
    sw, chunk = 480, 8
    cam_s, sh = scenes.camera_vectors(img_width=sw, aspect_ratio=ASPECT, **scenes._CAMERAS["vol2_final_scene"])
    S = O.OracleScene(cs)
    cam = O.make_camera(cam_s, BG, MAX_DEPTH)
    O.render(S, cam, sw, sh, 1, seed=SEED, mode=mode, nthreads=nthreads)            # thread-pool / cache warm-up
    done, t0 = 0, time.time()
    while time.time() - t0 < target_seconds:
        O.render(S, cam, sw, sh, chunk, seed=SEED, mode=mode, sample_begin=done, nthreads=nthreads)
        done += chunk
    dt = time.time() - t0
    return sw * sh * done / dt / 1e6, nthreads, f"vol2_final_scene {sw}x{sh} x {done} spp, depth {MAX_DEPTH} ({dt:.1f} s)"


def run_reference(args, rank):
    """--impl reference: the CPU arm.  Each step renders a bounded sample of the workload, on every core
    this process may use -- also under torchrun, which exports OMP_NUM_THREADS=1 to its workers."""
    if rank != 0:
        return
    import oracle as O
    from path_tracer_python_b200 import scenes
    O.build()
    nthreads = host_cores()
    cs = scenes.compile_world(scenes.vol2_final_scene(WIDTH, ASPECT, SPP, seed=SEED)[0])
    sw = 480
    cam_s, sh = scenes.camera_vectors(img_width=sw, aspect_ratio=ASPECT, **scenes._CAMERAS["vol2_final_scene"])
    S = O.OracleScene(cs)
    cam = O.make_camera(cam_s, BG, MAX_DEPTH)
    O.render(S, cam, sw, sh, 1, seed=SEED, nthreads=nthreads)    # thread pool start-up is not part of the probe
    t0 = time.time()
    O.render(S, cam, sw, sh, 2, seed=SEED, nthreads=nthreads)
    probe = (time.time() - t0) / 2
    budget = 150.0 / max(1, args.steps + args.warmup)            # whole run within a few minutes
    spp = max(1, min(4096, int(min(budget, 12.0) / max(probe, 1e-3))))
    for _ in range(args.warmup):
        O.render(S, cam, sw, sh, spp, seed=SEED, nthreads=nthreads)
    t0 = time.time()
    for _ in range(args.steps):
        O.render(S, cam, sw, sh, spp, seed=SEED, nthreads=nthreads)
    dt = (time.time() - t0) / max(1, args.steps)
    value = sw * sh * spp / dt / 1e6
    sample = f"vol2_final_scene {sw}x{sh} x {spp} spp per step, depth {MAX_DEPTH}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 and args.scaling != "weak" else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "vol2_final_scene 1920x1080 256spp depth50 (bounded CPU sample: " + sample + ")",
                   "mode": "megakernel semantics (trace_ray), CPU oracle port of the Taichi kernels",
                   "threads": nthreads, "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_RESULT_FD = None


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def kernel_stats(numerics):
    """Per-kernel instruction / DRAM / L1 counts of ONE config-2 step, from the committed ncu metrics pass
    of this workload (profiles/kernel_stats.json, written by tools/ncu_kernel_stats.py).  Counts of
    work done do not depend on the profiler; durations are never taken from it."""
    try:
        with open(os.path.join(REPO, "profiles", "kernel_stats.json")) as f:
            return json.load(f).get(numerics)
    except (OSError, ValueError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="wavefront", choices=["megakernel", "wavefront"])
    ap.add_argument("--scaling", default="auto", choices=["auto", "strong", "weak"])
    ap.add_argument("--numerics", default="strict", choices=["strict", "fast"])
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--spp", type=int, default=SPP)
    ap.add_argument("--min-batches", type=int, default=None, help="wavefront: at least this many batches per render call")
    ap.add_argument("--flight", type=int, default=None, help="wavefront: samples of every pixel in flight (all pipelines together)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other mode / numerics, probes and config 5")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="scheduling option for the engine (pt_set_option); never changes a result bit")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    # banner on the first collective), so everything except the result line goes to stderr.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from path_tracer_python_b200 import scenes
    from path_tracer_python_b200.engine import Engine, algorithmic_bytes
    from path_tracer_python_b200.renderer import TaichiRenderer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD
    dev = torch.device("cuda", local_rank)
    strong = world > 1 and args.scaling != "weak"

    def make_renderer(width, spp_total, numerics):
        # the scene goes through the package's own classes, compiler and native SAH builder, like a user's would
        # (scenes.vol2_final_scene follows src/scenes.py:1152-1244; seed 1234 = the reference's geometry)
        world, cam = scenes.vol2_final_scene(width, ASPECT, spp_total, seed=SEED)
        r = TaichiRenderer(world, cam, os.path.join("/tmp", f"bench_{rank}.png"), seed=SEED, device=local_rank,
                           process_group=group, quiet=True, numerics=numerics, reduce_dst=0)
        cs = r.compiled
        r.background_color, r.max_depth = BG, MAX_DEPTH
        r._upload_camera()
        for kv in args.opt:
            k, v = kv.split("=")
            r.engine.set_option(k, int(v))
        if args.min_batches:
            r.engine.wavefront_min_batches = args.min_batches
        if args.flight:
            r.engine.wavefront_samples_in_flight = args.flight
            r.engine.max_workspace_bytes = 80 << 30
        return r, cs

    # strong scaling: the frame's 256 samples are shared out; weak: every rank renders 256 of its own
    spp_total = args.spp if (strong or world == 1) else args.spp * world
    r, cs = make_renderer(args.width, spp_total, args.numerics)
    cam = r.cam
    W, H = cam.img_width, cam.img_height
    eng = r.engine
    render = eng.render_megakernel if args.mode == "megakernel" else eng.render_wavefront
    s_begin, s_end = r._sample_range()
    n_samples = s_end - s_begin
    paths_per_step = W * H * spp_total
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)        # > 126 MB L2

    def reduce_():
        if group is not None:
            dist.reduce(r._accum, dst=0, op=dist.ReduceOp.SUM, group=group)

    def step():
        r._accum.zero_()
        render(r._accum, s_begin, n_samples, SEED)
        reduce_()

    def barrier():
        if group is not None:
            dist.barrier(group=group)
        torch.cuda.synchronize(dev)

    def gmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if group is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())

    def gather_list(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if group is None:
            return [float(x)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t, group=group)
        return [float(o.item()) for o in out]

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: exactly K steps, L2 flushed between steps, device timers
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(args.steps)]
    barrier()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin.record()
    for k in range(args.steps):
        flush.fill_(k & 255)
        r._accum.zero_()
        ev[k][0].record()
        render(r._accum, s_begin, n_samples, SEED)
        ev[k][1].record()
        reduce_()
        ev[k][2].record()
    t_end.record()
    barrier()
    clocks = sampler.stop()
    total_ms = gmax(t_begin.elapsed_time(t_end))
    kernel_ms = [a.elapsed_time(b) for a, b, _ in ev]
    reduce_ms = [b.elapsed_time(c) for _, b, c in ev]
    ms_per_step = total_ms / args.steps
    value = paths_per_step / (ms_per_step * 1e-3) / 1e6
    rank_render_ms = gather_list(sum(kernel_ms) / len(kernel_ms))
    rank_reduce_ms = gather_list(sum(reduce_ms) / len(reduce_ms))
    reduced = r._accum.clone() if rank == 0 else None      # the last step's frame (rank 0 holds the sum)

    # ---- e2e: host buffers in/out through the renderer's public call
    barrier()
    for _ in range(2):
        r.render_frame(args.mode)
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_accum = r.render_frame(args.mode)
    if group is not None:
        dist.barrier(group=group)
    e2e_s = gmax((time.perf_counter() - t0) / e2e_steps)
    e2e_value = paths_per_step / e2e_s / 1e6
    mean_radiance = float(host_accum.mean() / cam.samples_per_pixel) if rank == 0 else None

    def time_one(fn, reps=1):
        """device ms of fn() (warm), max over ranks"""
        fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(3)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        barrier()
        return gmax(a.elapsed_time(b) / reps)

    extras = not args.no_extras
    modes = {args.mode: value}
    numerics_modes = {args.numerics: value}
    nccl_check = config5 = primary = tile_split = None
    if extras:
        # ---- the other architecture, for the record (BASELINE config 2 is "megakernel vs wavefront")
        other_mode = "megakernel" if args.mode == "wavefront" else "wavefront"
        other_render = eng.render_megakernel if other_mode == "megakernel" else eng.render_wavefront

        def other_step():
            r._accum.zero_()
            other_render(r._accum, s_begin, n_samples, SEED)
            reduce_()
        modes[other_mode] = paths_per_step / (time_one(other_step) * 1e-3) / 1e6

        # ---- N > 1: the reduced frame against the same frame rendered by one GPU alone
        if group is not None and rank == 0 and strong:
            alone = eng.new_accum(W, H)
            render(alone, 0, spp_total, SEED)
            torch.cuda.synchronize(dev)
            diff = (reduced - alone).abs()
            rel = diff / alone.abs().clamp_min(1e-6)
            nccl_check = {"what": f"NCCL-reduced {world}-GPU frame vs the same {spp_total} spp rendered by rank 0 alone",
                          "max_abs_diff": float(diff.max()), "max_rel_diff": float(rel.max()),
                          "mean_abs": float(alone.abs().mean()),
                          "allclose_rtol_1e-6": bool(torch.allclose(reduced, alone, rtol=1e-6, atol=1e-6 * float(alone.abs().mean()))),
                          "allclose_rtol_1e-5": bool(torch.allclose(reduced, alone, rtol=1e-5, atol=1e-5 * float(alone.abs().mean())))}
            del alone
        barrier()

        # ---- N > 1: the other sharding -- interleaved 8-row blocks per rank (one render call each), gather instead
        #      of reduce: each rank ships only its own 1/N of the frame and nothing is summed
        if group is not None and strong:
            from path_tracer_python_b200.distributed import gather_rows, interleaved_window
            ty0, ty1, tstride = interleaved_window(H, rank, world, 8)

            def tile_step():
                r._accum.zero_()
                eng.set_row_interleave(8, tstride)
                render(r._accum, 0, spp_total, SEED, y0=ty0, y1=ty1)
                eng.set_row_interleave(0)
                gather_rows(r._accum, H, group=group, block=8, dst=0)
            tile_step()
            barrier()
            tev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            flush.fill_(7)
            r._accum.zero_()
            tev[0].record()
            eng.set_row_interleave(8, tstride)
            render(r._accum, 0, spp_total, SEED, y0=ty0, y1=ty1)
            eng.set_row_interleave(0)
            tev[1].record()
            gather_rows(r._accum, H, group=group, block=8, dst=0)
            tev[2].record()
            barrier()
            tile_ms = gmax(tev[0].elapsed_time(tev[2]))
            tile_split = {"what": "interleaved 8-row blocks per rank (pt_set_row_interleave, one call), gather onto rank 0",
                          "value": paths_per_step / (tile_ms * 1e-3) / 1e6, "unit": "Mpaths/s", "ms": tile_ms, "steps": 1,
                          "rank_render_ms": gather_list(tev[0].elapsed_time(tev[1])),
                          "rank_gather_ms": gather_list(tev[1].elapsed_time(tev[2]))}
            if rank == 0:
                alone = eng.new_accum(W, H)
                render(alone, 0, spp_total, SEED)
                torch.cuda.synchronize(dev)
                tile_split["bit_identical_to_single_gpu"] = bool(torch.equal(r._accum, alone))
                del alone
            barrier()

    # ---- roofline of the dominant kernel (rank 0).  Algorithmic bytes come from the counting instantiation
    #      of the same kernels on the identical (seed, samples) launch; durations are CUDA-event times on
    #      the launching stream: one extra, identical step with every launch bracketed by events
    #      (pt_enable_stage_timing), summed over the launches of each kernel.
    roof = cpu = None
    launches_per_step = 1
    if rank == 0:
        peaks, which = measured_peaks()
        k_ms = sum(kernel_ms) / len(kernel_ms)
        scratch = eng.new_accum(W, H)
        stage_ms = stage_cnt = None
        if args.mode == "wavefront":
            eng.enable_stage_timing(True)
            flush.fill_(1)
            render(scratch, s_begin, n_samples, SEED)
            stage_ms = eng.stage_times()
            launches_per_step = len(stage_ms)
        eng.enable_counters(True)
        eng.counters(reset=True)
        render(scratch, s_begin, n_samples, SEED)
        if args.mode == "wavefront":
            stage_cnt = eng.stage_counters()
            eng.enable_stage_timing(False)
        cnt = eng.counters(reset=True)
        eng.enable_counters(False)
        alg = algorithmic_bytes(cnt)
        pipeline = {"kernel_ms": k_ms, "algorithmic_bytes": alg, "achieved_gbs": alg / (k_ms * 1e-3) / 1e9,
                    "counters": cnt, "mrays_per_s": cnt["rays"] / (k_ms * 1e-3) / 1e6,
                    "bytes_per_path": alg / max(1, W * H * n_samples)}
        sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
        num_sms = torch.cuda.get_device_properties(dev).multi_processor_count
        issue_peak = num_sms * 128 * sm_mhz * 1e6 / 1e12        # thread-instructions / s: SMs x 4 schedulers x 32 lanes x clock
        # measured gather roofs of the node-fetch stream (pt_probe_gather: 64-byte records, the traversal
        # kernels' fetch sequence, all lanes active): working set inside L1, = this scene's blob (L2), beyond L2
        scene_bytes = int(eng.info.total_bytes)
        probes = None
        if extras:
            probes = {"l1_random_gbs": eng.probe_gather(32 << 10, 0), "l1_coherent_gbs": eng.probe_gather(32 << 10, 1),
                      "l2_scene_random_gbs": eng.probe_gather(scene_bytes, 0), "l2_116mb_random_gbs": eng.probe_gather(116 << 20, 0),
                      "hbm_random_gbs": eng.probe_gather(4 << 30, 0, iters=16),
                      "what": "pt_probe_gather: 64-B record gathers (4 x LDG.128), every lane active, 4 records in flight per lane; "
                              "random = one record per lane, coherent = one per warp", "scene_bytes": scene_bytes}
        stats = kernel_stats(args.numerics)
        spp_scale = (n_samples / stats["spp"]) if stats else None
        if args.mode == "megakernel":
            ach = pipeline["achieved_gbs"]
            roof = {"bound": "l2", "achieved": ach, "peak": probes["l2_scene_random_gbs"] if probes else None, "unit": "GB/s",
                    "frac": ach / probes["l2_scene_random_gbs"] if probes else None, "traffic": None, "peak_source": "pt_probe_gather (this run)",
                    "kernel": "k_megakernel", "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg, "launches": 1}
        else:
            per = {}
            for (name, wave, ms), c in zip(stage_ms, stage_cnt):
                d = per.setdefault(name, {"ms": 0.0, "launches": 0, "bytes": 0, "rays": 0})
                d["ms"] += ms; d["launches"] += 1; d["bytes"] += algorithmic_bytes(c); d["rays"] += c["rays"]
            ext = per["extend"]
            ach = ext["bytes"] / (ext["ms"] * 1e-3) / 1e9
            stage_total = sum(d["ms"] for d in per.values())
            roofs = {}
            if probes:
                roofs["l1_gather_random"] = {"achieved": ach, "peak": probes["l1_random_gbs"], "unit": "GB/s", "frac": ach / probes["l1_random_gbs"]}
                roofs["l2_gather_random"] = {"achieved": ach, "peak": probes["l2_scene_random_gbs"], "unit": "GB/s", "frac": ach / probes["l2_scene_random_gbs"]}
            traffic = None
            if stats and "k_wf_extend" in stats["kernels"]:
                ks = stats["kernels"]["k_wf_extend"]
                traffic = ks["dram_bytes"] * spp_scale / ext["launches"]
                tinst = ks["thread_inst"] * spp_scale / (ext["ms"] * 1e-3) / 1e12
                roofs["issue"] = {"achieved": tinst, "peak": issue_peak, "unit": "Tinst/s", "frac": tinst / issue_peak,
                                  "what": "thread-instructions executed (ncu smsp__thread_inst_executed.sum of the same launches, "
                                          "profiles/kernel_stats.json) / live event time, vs SMs x 128 lanes x SM clock",
                                  "warp_issue_frac": ks["warp_inst"] * spp_scale / (ext["ms"] * 1e-3) / (issue_peak * 1e12 / 32),
                                  "lanes_per_warp_inst": ks["thread_inst"] / ks["warp_inst"], "sm_mhz": sm_mhz}
                roofs["l1_bytes"] = {"achieved": ks["l1tex_t_bytes"] * spp_scale / (ext["ms"] * 1e-3) / 1e9, "unit": "GB/s",
                                     "peak": num_sms * 128 * sm_mhz * 1e6 / 1e9, "what": "ncu l1tex__t_bytes.sum / live time vs SMs x 128 B/clk"}
                roofs["l1_bytes"]["frac"] = roofs["l1_bytes"]["achieved"] / roofs["l1_bytes"]["peak"]
                roofs["hbm"] = {"achieved": ks["dram_bytes"] * spp_scale / (ext["ms"] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s"}
                roofs["hbm"]["frac"] = roofs["hbm"]["achieved"] / roofs["hbm"]["peak"]
            # the roof the kernel sits under = the closest ceiling that really is one (frac <= 1)
            under = {k: v for k, v in roofs.items() if v["frac"] <= 1.0}
            bound = max(under, key=lambda k: under[k]["frac"]) if under else None
            head = roofs.get(bound, {"achieved": ach, "peak": None, "unit": "GB/s", "frac": None})
            roof = {"bound": bound, "achieved": head["achieved"], "peak": head["peak"], "unit": head["unit"], "frac": head["frac"],
                    "traffic": traffic, "peak_source": "issue: SM count x 128 lanes x sampled SM clock; gathers: pt_probe_gather in this run; hbm: " + which,
                    "kernel": "k_wf_extend", "launches": ext["launches"], "kernel_ms": ext["ms"] / ext["launches"],
                    "algorithmic_bytes_per_launch": ext["bytes"] / ext["launches"], "algorithmic_gbs": ach,
                    "share_of_step": ext["ms"] / stage_total, "grays_per_s": ext["rays"] / (ext["ms"] * 1e-3) / 1e9,
                    "roofs": roofs, "stages_ms_per_step": {k: round(v["ms"], 3) for k, v in per.items()},
                    "stage_sum_over_step_ms": stage_total / k_ms}
            # the HBM-facing kernel: its own measured DRAM bytes against the HBM roof
            if stats and "k_wf_shade" in stats["kernels"] and "shade" in per:
                ks = stats["kernels"]["k_wf_shade"]
                g = ks["dram_bytes"] * spp_scale / (per["shade"]["ms"] * 1e-3) / 1e9
                roof["shade"] = {"bound": "hbm", "achieved": g, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": g / peaks["hbm_gbs"],
                                 "ms_per_step": per["shade"]["ms"], "dram_bytes_per_step": ks["dram_bytes"] * spp_scale}
        roof["note"] = ("algorithmic = 32 B per child-box test + 16/48/64 B per sphere/triangle/quad test; those fetches are served by "
                        "L1/L2 (scene = 2.5 MB), DRAM only carries ray queues and hit records: see DESIGN.md section 6")
        roof["probes"] = probes
        roof["pipeline"] = pipeline
        if extras:
            # ---- primary-ray traversal rate (BASELINE.json metric: "primary-ray Mrays/s"): pt_first_hit over one
            #      camera ray per pixel, 8 different samples, device-timed
            o_d = [eng.generate_camera_rays(W, H, SEED, s) for s in range(8)]
            eng.first_hit(*o_d[0], want_pn=False)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.fill_(5)
            a.record()
            for o, d in o_d:
                eng.first_hit(o, d, want_pn=False)
            b.record()
            torch.cuda.synchronize(dev)
            primary = {"primary_mrays_s": 8 * W * H / (a.elapsed_time(b) * 1e-3) / 1e6, "rays": 8 * W * H,
                       "what": "pt_first_hit (closest hit only) over the camera rays of 8 samples of every pixel"}
            if args.mode == "wavefront" and "primary" in per:
                primary["primary_wave_mpaths_s"] = W * H * n_samples / (per["primary"]["ms"] * 1e-3) / 1e6
        if not args.no_cpu_baseline and world == 1:
            v, cores, sample = cpu_oracle_rate(cs)
            cpu = {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample}
    barrier()

    if extras and args.numerics == "strict" and eng.has_numerics("fast"):
        # ---- the production-numerics build (FMA slab test, approximate reciprocals) on the same step
        rf, _ = make_renderer(args.width, spp_total, "fast")
        rend_f = rf.engine.render_megakernel if args.mode == "megakernel" else rf.engine.render_wavefront

        def fast_step():
            rf._accum.zero_()
            rend_f(rf._accum, s_begin, n_samples, SEED)
            if group is not None:
                dist.reduce(rf._accum, dst=0, op=dist.ReduceOp.SUM, group=group)
        numerics_modes["fast"] = paths_per_step / (time_one(fast_step) * 1e-3) / 1e6
        del rf

    # ---- config 5: vol2_final_scene_comparison 3840x2160 x 1024 spp, same sharding, one timed step
    if extras and not args.no_config5:
        del flush
        r5, _ = make_renderer(C5_WIDTH, C5_SPP if (strong or world == 1) else C5_SPP * world, args.numerics)
        b5, e5 = r5._sample_range()
        rend5 = r5.engine.render_megakernel if args.mode == "megakernel" else r5.engine.render_wavefront
        rend5(r5._accum, b5, min(8, e5 - b5), SEED)           # warm-up (workspace allocation, caches)
        barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        r5._accum.zero_()
        evs[0].record()
        rend5(r5._accum, b5, e5 - b5, SEED)
        evs[1].record()
        if group is not None:
            dist.reduce(r5._accum, dst=0, op=dist.ReduceOp.SUM, group=group)
        evs[2].record()
        barrier()
        ms5 = gmax(evs[0].elapsed_time(evs[2]))
        paths5 = r5.cam.img_width * r5.cam.img_height * int(r5.cam.samples_per_pixel)
        config5 = {"workload": f"vol2_final_scene_comparison {r5.cam.img_width}x{r5.cam.img_height} {int(r5.cam.samples_per_pixel)}spp depth{MAX_DEPTH}",
                   "value": paths5 / (ms5 * 1e-3) / 1e6, "unit": "Mpaths/s", "ms": ms5, "steps": 1, "paths": paths5,
                   "samples_per_rank": e5 - b5, "rank_render_ms": gather_list(evs[0].elapsed_time(evs[1])),
                   "rank_reduce_ms": gather_list(evs[1].elapsed_time(evs[2])),
                   "reduce_bytes": int(r5._accum.numel() * 4) if group is not None else 0,
                   "mean_radiance": float(r5._accum.mean() / r5.cam.samples_per_pixel) if rank == 0 else None}
        del r5

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"vol2_final_scene {W}x{H} {spp_total}spp depth{MAX_DEPTH}" + (f" ({n_samples} spp per GPU)" if world > 1 else ""),
                       "mode": args.mode,
                       "scene": "1007 spheres + 2401 quads, 6815-node SAH BVH (own scene classes -> compile_scene -> native SAH builder, seed 1234)",
                       "setup_ms": {k: round(v * 1e3, 2) for k, v in r.timing.items() if k in ("scene_compile", "bvh_compile", "gpu_upload", "total_setup")},
                       "l2": "flushed between timed steps (256 MiB fill)",
                       "parallelism": f"spp-split x{world}" + (", one NCCL reduce(sum) of the fp32 accumulation buffer onto rank 0 per step" if world > 1 else ""),
                       "numerics": "PT-F32 strict (bit-exact vs reference kernels)" if args.numerics == "strict" else
                                   "PT-F32 fast (FMA + approximate reciprocal; |dt|/t <= 1e-5 vs strict)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": r.frame_h2d_bytes,
                    "d2h_bytes_per_step": r.frame_d2h_bytes, "ms_per_step": e2e_s * 1e3},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof, "cpu_baseline": cpu, "mean_radiance": mean_radiance,
            "modes": modes, "numerics_modes": numerics_modes, "primary": primary,
            "ranks": {"render_ms": rank_render_ms, "reduce_ms": rank_reduce_ms,
                      "reduce_bytes": int(r._accum.numel() * 4) if world > 1 else 0, "samples_per_rank": n_samples},
            "nccl_vs_single_gpu": nccl_check, "tile_split": tile_split, "config5": config5,
        }
        emit(line)
    if group is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
