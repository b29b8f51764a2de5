"""Tuning sweep for the wavefront pipeline on config 2 (development aid)."""
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
def run():
    acc.zero_(); e.render_wavefront(acc, 0, SPP, 7); torch.cuda.synchronize()   # same size as the timed call (workspace allocation)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc.zero_(); a.record(); e.render_wavefront(acc, 0, SPP, 7); b.record(); torch.cuda.synchronize()
    chk = int(acc.view(torch.int32).to(torch.int64).sum().item()) & 0xffffffff    # bit-level checksum: must not depend on the options
    return W * H * SPP / a.elapsed_time(b) / 1e3, chk
L = lambda k, d: [int(x) for x in os.environ.get(k, d).split(",")]
for fl, bn, bps, sn, so, vn, wv, fu, mb, tl in itertools.product(L("FLIGHT", "32"), L("BURSTS", "20"), L("BPS", "7"), L("SNODES", "256"), L("SORT", "1"), L("VOTES", "16"), L("WAVES", "8"), L("FUSE", "1"), L("MINB", "5"), L("TILE", "1")):
    e.wavefront_samples_in_flight = fl
    e.set_option("wf_burst_num", bn); e.set_option("wf_blocks_per_sm", bps); e.set_option("smem_nodes", sn); e.set_option("wavefront_sort", so); e.set_option("wf_vote_num", vn); e.set_option("wf_waves", wv); e.set_option("wf_fuse_primary", fu); e.set_option("wf_extend_minb", mb); e.set_option("wf_tile", tl)
    v, chk = run()
    print(f"in_flight={fl} wf_burst={bn} blocks/sm={bps} smem_nodes={sn} sort={so} vote={vn} waves={wv} fuse={fu} minb={mb} tile={tl}: {v:.1f} Mpaths/s  chk={chk:08x}", flush=True)
