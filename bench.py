#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 path-tracing hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode megakernel|wavefront]

Workload (BASELINE.json configs[1]): vol2_final_scene (1007 spheres, 2401 quads, 6815-node SAH BVH,
fog dome, Perlin + image textures), 1920x1080, 256 spp, max_depth 50, black background.
A *step* = one full render pass: W*H*256 paths through the megakernel (or the wavefront pipeline).

  value  : Mpaths/s, whole job, scene blob + accumulation buffer resident in HBM, CUDA events.
  e2e    : Mpaths/s through the reference-facing API TaichiRenderer.render_frame(): packed scene
           copied pinned-host -> device, camera upload, render, accumulation buffer copied back
           to pinned host memory, all inside the timed region.
  N > 1  : one process per GPU (torchrun); rank r renders its own 256 samples [r*256,(r+1)*256)
           of the same frame (weak scaling: per-GPU work fixed) and the fp32 accumulation
           buffers are summed with one NCCL all-reduce inside the timed region.
  --impl reference : the CPU oracle (a scalar C++/OpenMP restatement of the reference's Taichi
           kernels -- Taichi itself is not installable here) on all host threads, on a bounded
           sample of the same workload.
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


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


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
    sw, chunk = 480, 8
    cam_s, sh = scenes.camera_vectors(img_width=sw, aspect_ratio=ASPECT, **scenes._CAMERAS["vol2_final_scene"])
    S = O.OracleScene(cs)
    cam = O.make_camera(cam_s, BG, MAX_DEPTH)
    O.render(S, cam, sw, sh, 1, seed=SEED, mode=mode)            # thread-pool / cache warm-up
    done, t0 = 0, time.time()
    while time.time() - t0 < target_seconds:
        O.render(S, cam, sw, sh, chunk, seed=SEED, mode=mode, sample_begin=done)
        done += chunk
    dt = time.time() - t0
    return sw * sh * done / dt / 1e6, O.num_threads(), f"vol2_final_scene {sw}x{sh} x {done} spp, depth {MAX_DEPTH} ({dt:.1f} s)"


def run_reference(args, rank):
    """--impl reference: the CPU arm.  Each step renders a bounded sample of the workload."""
    if rank != 0:
        return
    import oracle as O
    from path_tracer_python_b200 import scenes
    O.build()
    cs, meta = scenes.load_config("vol2_final_scene", WIDTH, ASPECT)
    sw = 480
    cam_s, sh = scenes.camera_vectors(img_width=sw, aspect_ratio=ASPECT, **scenes._CAMERAS["vol2_final_scene"])
    S = O.OracleScene(cs)
    cam = O.make_camera(cam_s, BG, MAX_DEPTH)
    t0 = time.time()
    O.render(S, cam, sw, sh, 1, seed=SEED)
    probe = time.time() - t0
    budget = 150.0 / max(1, args.steps + args.warmup)            # whole run within a few minutes
    spp = max(1, min(4096, int(min(budget, 12.0) / max(probe, 1e-3))))
    for _ in range(args.warmup):
        O.render(S, cam, sw, sh, spp, seed=SEED)
    t0 = time.time()
    for _ in range(args.steps):
        O.render(S, cam, sw, sh, spp, seed=SEED)
    dt = (time.time() - t0) / max(1, args.steps)
    value = sw * sh * spp / dt / 1e6
    cores = O.num_threads()
    sample = f"vol2_final_scene {sw}x{sh} x {spp} spp per step, depth {MAX_DEPTH}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "vol2_final_scene 1920x1080 256spp depth50 (bounded CPU sample: " + sample + ")",
                   "mode": "megakernel semantics (trace_ray), CPU oracle port of the Taichi kernels"},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample},
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


def recorded_traffic(mode):
    """dram__bytes_read + dram__bytes_write per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this workload (profiles/traffic.json; written by tools/ncu_traffic.py);
    None when no capture is recorded."""
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(mode, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="wavefront", choices=["megakernel", "wavefront"])
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--spp", type=int, default=SPP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    import torch
    import torch.distributed as dist
    from path_tracer_python_b200 import scenes
    from path_tracer_python_b200.engine import algorithmic_bytes
    from path_tracer_python_b200.renderer import TaichiRenderer
    from path_tracer_python_b200.scene import camera, point3, vec3

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD
    dev = torch.device("cuda", local_rank)

    cs, meta = scenes.load_config("vol2_final_scene", args.width, ASPECT)
    cam = camera()
    cam.aspect_ratio, cam.img_width = ASPECT, args.width
    cam.samples_per_pixel = args.spp * world          # weak scaling: every rank renders args.spp samples
    c = scenes._CAMERAS["vol2_final_scene"]
    cam.vfov, cam.lookfrom, cam.lookat, cam.vup = c["vfov"], point3(*c["lookfrom"]), point3(*c["lookat"]), vec3(*c["vup"])
    r = TaichiRenderer(None, cam, os.path.join("/tmp", f"bench_{rank}.png"), seed=SEED, device=local_rank,
                       process_group=group, compiled_scene=cs, quiet=True)
    r.background_color, r.max_depth = BG, MAX_DEPTH
    r._upload_camera()
    W, H = cam.img_width, cam.img_height
    eng = r.engine
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    render = eng.render_megakernel if args.mode == "megakernel" else eng.render_wavefront
    s_begin, s_end = r._sample_range()
    n_samples = s_end - s_begin
    paths_per_step = W * H * n_samples * world
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)        # > 126 MB L2

    def step():
        r._accum.zero_()
        render(r._accum, s_begin, n_samples, SEED)
        if group is not None:
            dist.all_reduce(r._accum, group=group)

    def barrier():
        if group is not None:
            dist.barrier(group=group)
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: exactly K steps, L2 flushed between steps, device timers
    sampler = ClockSampler(local_rank)
    sampler.start()
    kernel_ms = []
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin.record()
    for k in range(args.steps):
        flush.fill_(k & 255)
        r._accum.zero_()
        ev[k][0].record()
        render(r._accum, s_begin, n_samples, SEED)
        ev[k][1].record()
        if group is not None:
            dist.all_reduce(r._accum, group=group)
    t_end.record()
    barrier()
    clocks = sampler.stop()
    total_ms = t_begin.elapsed_time(t_end)
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = paths_per_step / (ms_per_step * 1e-3) / 1e6

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
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    e2e_s = float(t.item())
    e2e_value = paths_per_step / e2e_s / 1e6
    mean_radiance = float(host_accum.mean() / cam.samples_per_pixel)

    # ---- the other architecture, for the record (BASELINE config 2 is "megakernel vs wavefront")
    other_mode = "megakernel" if args.mode == "wavefront" else "wavefront"
    other_render = eng.render_megakernel if other_mode == "megakernel" else eng.render_wavefront
    barrier()
    r._accum.zero_()
    other_render(r._accum, s_begin, n_samples, SEED)
    barrier()
    oa, ob = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r._accum.zero_()
    oa.record()
    other_render(r._accum, s_begin, n_samples, SEED)
    ob.record()
    barrier()
    t = torch.tensor([oa.elapsed_time(ob)], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    other_value = paths_per_step / (float(t.item()) * 1e-3) / 1e6

    # ---- roofline of the dominant kernel.  Algorithmic bytes come from the counting instantiation of
    #      the same kernels on the identical (seed, samples) launch; durations are CUDA-event times
    #      on the launching stream.  Megakernel: the one launch of the timed steps above.  Wavefront:
    #      one extra, identical step with every launch bracketed by events (pt_enable_stage_timing),
    #      summed over the launches of k_wf_extend (the kernel that owns >50 % of a step).
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
        traffic = recorded_traffic(args.mode)
        if args.mode == "megakernel":
            roof = {"bound": "hbm", "achieved": pipeline["achieved_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": pipeline["achieved_gbs"] / peaks["hbm_gbs"], "traffic": traffic, "peak_source": which,
                    "kernel": "k_megakernel", "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg, "launches": 1}
        else:
            per = {}
            for (name, wave, ms), c in zip(stage_ms, stage_cnt):
                d = per.setdefault(name, {"ms": 0.0, "launches": 0, "bytes": 0, "rays": 0})
                d["ms"] += ms; d["launches"] += 1; d["bytes"] += algorithmic_bytes(c); d["rays"] += c["rays"]
            ext = per["extend"]
            ach = ext["bytes"] / (ext["ms"] * 1e-3) / 1e9
            stage_total = sum(d["ms"] for d in per.values())
            roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                    "traffic": traffic, "peak_source": which, "kernel": "k_wf_extend",
                    "launches": ext["launches"], "kernel_ms": ext["ms"] / ext["launches"],
                    "algorithmic_bytes_per_launch": ext["bytes"] / ext["launches"],
                    "share_of_step": ext["ms"] / stage_total, "grays_per_s": ext["rays"] / (ext["ms"] * 1e-3) / 1e9,
                    "stages_ms_per_step": {k: round(v["ms"], 3) for k, v in per.items()},
                    "stage_sum_over_step_ms": stage_total / k_ms}
        roof["note"] = ("algorithmic = 32 B per child-box test + 16/48/64 B per sphere/triangle/quad test; those fetches are "
                        "served by shared memory + L1/L2 (scene = 2.5 MB), so DRAM traffic << algorithmic: see profiles/")
        roof["pipeline"] = pipeline
        if not args.no_cpu_baseline and world == 1:
            v, cores, sample = cpu_oracle_rate(cs)
            cpu = {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"vol2_final_scene {W}x{H} {n_samples}spp/GPU depth{MAX_DEPTH}", "mode": args.mode,
                       "scene": "1007 spheres + 2401 quads, 6815-node SAH BVH (reference compile, seed 1234)",
                       "l2": "flushed between timed steps (256 MiB fill)", "parallelism": f"spp-split x{world}",
                       "numerics": "PT-F32 strict (bit-exact vs reference kernels)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": r.frame_h2d_bytes,
                    "d2h_bytes_per_step": r.frame_d2h_bytes, "ms_per_step": e2e_s * 1e3},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof, "cpu_baseline": cpu, "mean_radiance": mean_radiance,
            "modes": {args.mode: value, other_mode: other_value},
        }
        emit(line)
    if group is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
