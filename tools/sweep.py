"""Tuning sweep (development aid): megakernel throughput on config 2 vs burst_num / smem_nodes / block size."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))   # golden_fixtures
import torch
from path_tracer_python_b200.engine import Engine
from path_tracer_python_b200 import scenes
import golden_fixtures
W = int(os.environ.get("W", 1920)); SPP = int(os.environ.get("SPP", 32))
e = Engine(0)
cs, meta = golden_fixtures.load_config("vol2_final_scene", W, 16 / 9)
e.set_scene(cs); e.set_camera(meta["camera"], (0, 0, 0), 50)
H = meta["height"]
acc = e.new_accum(W, H)
def run(mode="megakernel"):
    global SPP
    fn = e.render_megakernel if mode == "megakernel" else e.render_wavefront
    acc.zero_(); fn(acc, 0, 2, 7); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc.zero_(); a.record(); fn(acc, 0, SPP, 7); b.record(); torch.cuda.synchronize()
    return W * H * SPP / a.elapsed_time(b) / 1e3
bursts = [int(x) for x in os.environ.get("BURSTS", "0,8,12,16,20,24,28").split(",")]
snodes = [int(x) for x in os.environ.get("SNODES", "1024").split(",")]
blocks = [int(x) for x in os.environ.get("BLOCKS", "256").split(",")]
syncs = [int(x) for x in os.environ.get("SYNCS", "0").split(",")]
spps = [int(x) for x in os.environ.get("SPPS", str(SPP)).split(",")]
votes = [int(x) for x in os.environ.get("VOTES", "16").split(",")]
for bt, sn, bn, sy, sp, vn in itertools.product(blocks, snodes, bursts, syncs, spps, votes):
    e.set_option("block_threads", bt); e.set_option("smem_nodes", sn); e.set_option("burst_num", bn); e.set_option("sync_samples", sy); e.set_option("vote_num", vn)
    SPP = sp
    print(f"block={bt} smem_nodes={sn} burst={bn} sync={sy} spp={sp} vote={vn}: {run():.1f} Mpaths/s", flush=True)
if os.environ.get("WAVE"):
    print(f"wavefront: {run('wavefront'):.1f} Mpaths/s")
