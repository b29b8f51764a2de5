"""Join an ncu SASS-level source page with nvdisasm line info and aggregate by CUDA source line.

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top_n]
Prints, per source line: warp-level instructions executed, share, and average active threads.
"""
import csv, os, re, subprocess, sys, collections, tempfile
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(REPO, "path_tracer_python_b200", "libb200pt.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# parse nvdisasm: function sections ".text.<name>:", line markers "//## File "...", line N", instructions "/*addr*/ OP ..."
funcs = {}
cur = None; line = None; fname = None
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        cur = m.group(1); funcs[cur] = []; line = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        fname, line = os.path.basename(m.group(1)), int(m.group(2)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), fname, line, m.group(2)))
fn = [k for k in funcs if re.search(kre, k)]
assert fn, list(funcs)
instrs = funcs[fn[0]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
print(f"kernel {fn[0]}: {len(instrs)} SASS instrs (nvdisasm) vs {len(data)} (ncu)")
n = min(len(instrs), len(data))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
tot = 0.0
for (addr, f, ln, txt), r in zip(instrs[:n], data[:n]):
    w = float(r[ix["Instructions Executed"]] or 0); t = float(r[ix["Thread Instructions Executed"]] or 0)
    smp = float(r[ix["# Samples"]] or 0)
    a = agg[(f, ln)]; a[0] += w; a[1] += t; a[2] += smp; tot += w
print(f"total warp instr {tot:.3e}")
src_cache = {}
def src(f, ln):
    for d in ("path_tracer_python_b200/csrc",):
        p = os.path.join(REPO, d, f or "")
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][ln - 1].strip()[:90] if ln and ln <= len(src_cache[p]) else ""
    return ""
tsmp = sum(a[2] for a in agg.values())
for (f, ln), (w, t, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{w/tot*100:5.1f}% inst  {smp/max(tsmp,1)*100:5.1f}% samples  avg_thr {t/max(w,1):5.1f}  {f}:{ln}  {src(f, ln)}")
