"""N-GPU parity: the NCCL-reduced frame of a 2-rank spp split equals the frame one GPU renders alone
(same sample set; only the fp32 summation order differs).  Needs >= 2 devices (skipped on 1-GPU boxes;
bench.py reports the same comparison at every N > 1 as `nccl_vs_single_gpu`)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
sys.path.insert(0, %(repo)r); sys.path.insert(0, os.path.join(%(repo)r, "tests"))
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from path_tracer_python_b200 import scenes
import golden_fixtures
from path_tracer_python_b200.engine import Engine
from path_tracer_python_b200.distributed import sample_slice, reduce_accum, interleaved_window, gather_rows
cs, meta = golden_fixtures.load_config("vol2_final_scene", 480, 16 / 9)
W, H, spp, seed = 480, meta["height"], 64, 5
e = Engine(rank)
e.set_scene(cs); e.set_camera(meta["camera"], (0, 0, 0), 50)
ok = True
for mode in ("wavefront", "megakernel"):
    fn = e.render_wavefront if mode == "wavefront" else e.render_megakernel
    # spp split + reduce onto rank 0
    b, en = sample_slice(spp)
    part = e.new_accum(W, H); fn(part, b, en - b, seed)
    reduce_accum(part, dst=0)
    # row-block (tile) split + gather: disjoint rows, no summation at all -> bit-identical
    tile = e.new_accum(W, H)
    y0, y1, stride = interleaved_window(H, rank, world, 8)
    e.set_row_interleave(8, stride)
    fn(tile, 0, spp, seed, y0=y0, y1=y1)               # this rank's interleaved 8-row blocks, one call
    e.set_row_interleave(0)
    gather_rows(tile, H, block=8, dst=0)
    if rank == 0:
        alone = e.new_accum(W, H); fn(alone, 0, spp, seed)
        d = (part - alone).abs().max().item()
        close = torch.allclose(part, alone, rtol=1e-5, atol=1e-5 * alone.abs().mean().item())
        same = torch.equal(tile, alone)
        print(f"{mode}: spp-split max abs diff {d:.3e} allclose={close}; tile-split bit-identical={same}")
        ok = ok and close and same
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


@pytest.mark.timeout(600)
def test_two_gpu_reduce_equals_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"repo": REPO})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("bit-identical=True") == 2 and r.stdout.count("allclose=True") == 2
