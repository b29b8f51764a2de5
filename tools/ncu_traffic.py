"""Record the measured DRAM traffic of the dominant kernel for bench.py's `roofline.traffic`.

    python tools/ncu_traffic.py <mode> <report.ncu-rep> <kernel-regex> [command that was profiled]

Reads dram__bytes_read.sum + dram__bytes_write.sum of every launch in the report whose name matches
(one `ncu --set full` capture of the bench workload), averages per launch, and updates
profiles/traffic.json[mode].  bench.py only reads that file; a number taken under the profiler is
never a bench value, but bytes moved do not depend on the profiler."""
import csv, json, os, re, subprocess, sys
mode, rep, kre = sys.argv[1], sys.argv[2], sys.argv[3]
cmd = sys.argv[4] if len(sys.argv) > 4 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
per = []
for vals in rows[2:]:
    if not re.search(kre, vals[ix["Kernel Name"]]):
        continue
    rd = to_bytes(vals[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(vals[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    ms = float(vals[ix["gpu__time_duration.sum"]]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[ix["gpu__time_duration.sum"]]]
    per.append({"read": rd, "write": wr, "ms_under_ncu": ms})
assert per, "no matching launches"
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[mode] = {"kernel_regex": kre, "launches": len(per), "report": os.path.basename(rep), "command": cmd,
              "dram_bytes_per_launch": sum(p["read"] + p["write"] for p in per) / len(per),
              "dram_read_bytes_per_launch": sum(p["read"] for p in per) / len(per),
              "dram_write_bytes_per_launch": sum(p["write"] for p in per) / len(per),
              "per_launch": per}
json.dump(data, open(path, "w"), indent=1)
print(json.dumps({k: v for k, v in data[mode].items() if k != "per_launch"}))
