"""A/B timing of library variants on config 2 (development aid).

    python tools/ab.py libA.so libB.so ...      (paths relative to path_tracer_python_b200/)

Each variant runs in its own process (the library is loaded once per process), variants are
interleaved over REPS rounds so that clock / box drift hits all of them alike; prints the best and
median Mpaths/s of both architectures and a bit-level checksum (must be identical across variants).
"""
import os, subprocess, sys, statistics
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
if os.environ.get("AB_CHILD"):
    sys.path.insert(0, ROOT)
    import torch
    from path_tracer_python_b200.engine import Engine
    from path_tracer_python_b200 import scenes
    sys.path.insert(0, HERE)
    import _scenes
    SPP = int(os.environ.get("SPP", 64))
    e = Engine(0)
    cs, meta, W = _scenes.load()
    e.set_scene(cs); e.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    H = meta["height"]
    if os.environ.get("FLIGHT"):
        e.wavefront_samples_in_flight = int(os.environ["FLIGHT"])
    for k, v in [kv.split("=") for kv in os.environ.get("OPTS", "").split(",") if kv]:
        e.set_option(k, int(v))
    out = []
    for mode in os.environ.get("MODES", "wavefront,megakernel").split(","):
        fn = e.render_megakernel if mode == "megakernel" else e.render_wavefront
        acc = e.new_accum(W, H)
        fn(acc, 0, SPP, 7); torch.cuda.synchronize()
        best = 0.0
        for _ in range(2):
            acc.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(acc, 0, SPP, 7); b.record(); torch.cuda.synchronize()
            best = max(best, W * H * SPP / a.elapsed_time(b) / 1e3)
        chk = int(acc.view(torch.int32).to(torch.int64).sum().item()) & 0xffffffff
        out.append(f"{mode}={best:.1f}:{chk:08x}")
    print("AB " + " ".join(out))
    sys.exit(0)
libs = sys.argv[1:]
reps = int(os.environ.get("REPS", 3))
res = {l: {} for l in libs}
for r in range(reps):
    for l in libs:
        env = dict(os.environ, AB_CHILD="1", PT_LIB_PATH=os.path.join(ROOT, "path_tracer_python_b200", l))
        p = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=600)
        line = [x for x in p.stdout.splitlines() if x.startswith("AB ")]
        if not line:
            print(l, "FAILED", p.stderr[-400:]); continue
        for tok in line[0].split()[1:]:
            mode, rest = tok.split("="); v, chk = rest.split(":")
            res[l].setdefault(mode, []).append((float(v), chk))
for l in libs:
    for mode, vals in res[l].items():
        v = [x[0] for x in vals]
        print(f"{l:34s} {mode:11s} best {max(v):7.1f}  median {statistics.median(v):7.1f}  runs {['%.0f' % x for x in v]}  chk {sorted(set(x[1] for x in vals))}", flush=True)
