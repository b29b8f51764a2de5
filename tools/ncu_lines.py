"""Per-source-line (and per-function) instruction / stall-sample shares of one kernel in an ncu report
captured with `--set full --import-source on` from a `-lineinfo` build.

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top_n] [--funcs]

Uses ncu's own correlated view (`--page source --print-source cuda,sass`), so the source text is
the one embedded in the report -- it stays valid after the tree has moved on.  `--funcs` aggregates
the lines by the enclosing function of the embedded source instead of listing single lines.
"""
import collections, csv, re, subprocess, sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
rep, kre = args[0], args[1]
topn = int(args[2]) if len(args) > 2 else 40
by_func = "--funcs" in sys.argv

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines = []            # (file, line_no, text, warp_instr, thread_instr, samples)
sources = collections.defaultdict(dict)
fpath = func = None
ix = None
seen_funcs = []
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        func = rows[i + 1][1]
        hdr = rows[i + 2]
        ix = {}
        for k, h in enumerate(hdr):
            ix.setdefault(h, k)
        if func not in seen_funcs:
            seen_funcs.append(func)
        i += 3
        continue
    if ix and r and r[0] not in ("", "Line No") and re.search(kre, func or ""):
        try:
            ln = int(r[0])
        except ValueError:
            i += 1
            continue
        def num(name):
            v = r[ix[name]] if name in ix and ix[name] < len(r) else "0"
            try:
                return float(v)
            except ValueError:
                return 0.0
        lines.append((fpath, ln, r[1].strip(), num("Instructions Executed"), num("Thread Instructions Executed"),
                      num("# Samples")))
    elif ix and r and r[0] not in ("", "Line No"):
        pass
    if ix and r and r[0].isdigit():
        sources[fpath][int(r[0])] = r[1]
    i += 1
if not lines:
    sys.exit(f"no kernel matches {kre!r}; kernels in the report: {seen_funcs}")
tot = sum(l[3] for l in lines)
tsmp = sum(l[5] for l in lines) or 1.0
tthr = sum(l[4] for l in lines)
print(f"kernel /{kre}/: {tot:.3e} warp instructions, {tthr / max(tot, 1):.1f} active threads on average")

if by_func:
    # full embedded sources (the correlated view only lists lines that own instructions)
    full = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda", "--csv"],
                          capture_output=True, text=True).stdout
    cur = None
    for r in csv.reader(full.splitlines()):
        if r and r[0] in ("File Name", "File Path"):
            cur = r[1].split("/")[-1]
        elif cur and r and r[0].isdigit():
            sources[cur][int(r[0])] = r[1]
    # headers are not embedded in that view: fall back to the working tree (exact for a fresh profile)
    import os
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "path_tracer_python_b200", "csrc")
    for f in {l[0] for l in lines}:
        pth = os.path.join(csrc, f)
        if f != "b200pt.cu" and os.path.exists(pth):
            for k, t in enumerate(open(pth).read().splitlines(), 1):
                sources[f][k] = t
    # enclosing function = the nearest preceding line of the same file that looks like a definition
    fdef = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__)[^;(]*?\b(\w+)\s*\((?!\d)")
    cache = {}
    def enclosing(f, ln):
        src = sources.get(f, {})
        for k in range(ln, 0, -1):
            t = src.get(k)
            if t is None:
                continue
            m = fdef.match(re.sub(r"__launch_bounds__\([^)]*\)", "", t.strip()))
            if m:
                return f"{f}:{m.group(1)}"
        return f
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    for f, ln, text, w, t, s in lines:
        a = agg[enclosing(f, ln)]
        a[0] += w; a[1] += t; a[2] += s
    for k, (w, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        print(f"{w / tot * 100:5.1f}% inst  {s / tsmp * 100:5.1f}% samples  avg_thr {t / max(w, 1):5.1f}  {k}")
else:
    for f, ln, text, w, t, s in sorted(lines, key=lambda l: -l[3])[:topn]:
        print(f"{w / tot * 100:5.1f}% inst  {s / tsmp * 100:5.1f}% samples  avg_thr {t / max(w, 1):5.1f}  {f}:{ln}  {text[:90]}")
