"""Measured roofs of the node-fetch stream (pt_probe_gather) over working sets from inside L1 to beyond L2
-> profiles/<tag>_gather_peaks.json.      python tools/gather_peaks.py r02"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from path_tracer_python_b200.engine import Engine
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
e = Engine(0)
sizes = [16 << 10, 32 << 10, 64 << 10, 128 << 10, 256 << 10, 1 << 20, 2504448, 8 << 20, 32 << 20, 64 << 20, 116 << 20,
         256 << 20, 1 << 30, 4 << 30]
out = {"device": torch.cuda.get_device_name(0), "what": "pt_probe_gather: 64-byte record gathers (4 x LDG.128 through the read-only path), "
       "192-thread CTAs x 6 per SM, every lane active, 4 records in flight per lane; GB/s of gathered bytes, best of 5 launches",
       "random_per_lane": {}, "one_record_per_warp": {}}
for s in sizes:
    it = 64 if s <= (256 << 20) else 16
    out["random_per_lane"][str(s)] = e.probe_gather(s, 0, iters=it)
    out["one_record_per_warp"][str(s)] = e.probe_gather(s, 1, iters=it)
    print(f"{s / 1024:12.0f} KiB  random {out['random_per_lane'][str(s)]:9.1f} GB/s   coherent {out['one_record_per_warp'][str(s)]:9.1f} GB/s", flush=True)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"{tag}_gather_peaks.json")
json.dump(out, open(path, "w"), indent=1)
print(path)
