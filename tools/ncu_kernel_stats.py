"""Per-kernel work counts of one config-2 batch from an ncu metrics pass -> profiles/kernel_stats.json,
which bench.py combines with its own live event timing into the issue / L1 / HBM roofs.

    ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,\
dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_bytes.sum,lts__t_bytes.sum \
        --clock-control none --csv --log-file gpurun_out/r02_stats.csv env SPP=32 MODES=wavefront python tools/profile_run.py
    python tools/ncu_kernel_stats.py strict gpurun_out/r02_stats.csv 32 "<the command above>"

Counts of work done (instructions, bytes) do not depend on the profiler; durations are never taken
from it (they are recorded only so that the kernel SHARES can be cross-checked against live timing).
profile_run.py renders 2 one-sample warm-up calls and then ONE call of SPP samples: only the launches of
that last call are counted (everything after the last k_wf_reset / before it for the megakernel)."""
import csv, json, os, re, sys

numerics, path, spp = sys.argv[1], sys.argv[2], int(sys.argv[3])
cmd = sys.argv[4] if len(sys.argv) > 4 else ""
rows = [r for r in csv.reader(open(path, errors="replace")) if r and not r[0].startswith("==")]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "inst": 1, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9,
         "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9, "": 1}
launches = {}          # id -> {kernel, metrics}
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    lid = int(r[ix["ID"]])
    d = launches.setdefault(lid, {"kernel": r[ix["Kernel Name"]], "m": {}})
    unit = r[ix["Metric Unit"]]
    d["m"][r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * scale.get(unit, 1)
ids = sorted(launches)
# profile_run.py makes 2 one-sample warm-up calls (one batch = one k_wf_reset each), then the measured call of SPP
# samples (one reset per batch, however many batches the engine splits it into): count from the third reset on
resets = [i for i in ids if "k_wf_reset" in launches[i]["kernel"]]
first = resets[2] if len(resets) > 2 else (resets[-1] if resets else ids[0])
per = {}
for i in ids:
    if i < first:
        continue
    k = re.sub(r"<.*", "", launches[i]["kernel"].split("(")[0]).replace("void ", "").strip()
    m = launches[i]["m"]
    d = per.setdefault(k, {"launches": 0, "thread_inst": 0.0, "warp_inst": 0.0, "dram_bytes": 0.0, "l1tex_t_bytes": 0.0,
                           "lts_t_bytes": 0.0, "ns_under_ncu": 0.0})
    d["launches"] += 1
    d["thread_inst"] += m.get("smsp__thread_inst_executed.sum", 0.0)
    d["warp_inst"] += m.get("smsp__inst_executed.sum", 0.0)
    d["dram_bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    d["l1tex_t_bytes"] += m.get("l1tex__t_bytes.sum", 0.0)
    d["lts_t_bytes"] += m.get("lts__t_bytes.sum", 0.0)
    d["ns_under_ncu"] += m.get("gpu__time_duration.sum", 0.0)
out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "kernel_stats.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
data[numerics] = {"spp": spp, "workload": "vol2_final_scene 1920x1080, one render call of `spp` samples (wavefront)",
                  "source": os.path.basename(path), "command": cmd, "kernels": per}
json.dump(data, open(out_path, "w"), indent=1)
tot = sum(d["ns_under_ncu"] for d in per.values())
for k, d in sorted(per.items(), key=lambda kv: -kv[1]["ns_under_ncu"]):
    print(f"{k:22s} x{d['launches']:3d}  {d['ns_under_ncu'] / 1e6:8.3f} ms ({100 * d['ns_under_ncu'] / tot:4.1f} %)  "
          f"thread-inst {d['thread_inst']:.3e}  lanes/warp-inst {d['thread_inst'] / max(1, d['warp_inst']):5.2f}  "
          f"dram {d['dram_bytes'] / 1e9:7.3f} GB  l1tex {d['l1tex_t_bytes'] / 1e9:8.2f} GB  l2 {d['lts_t_bytes'] / 1e9:8.2f} GB")
