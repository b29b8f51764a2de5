"""N > 1 host logic on CPU: world_size-2 gloo run of the sample split + accumulation-buffer sum."""
import os
import socket
import subprocess
import sys

import numpy as np

from path_tracer_python_b200.distributed import interleaved_window, row_blocks, row_slice, sample_slice

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_slices_partition_every_spp():
    for world in (1, 2, 3, 4, 8):
        for spp in (0, 1, 5, 16, 255, 256, 1024):
            seen = []
            for r in range(world):
                b, e = sample_slice(spp, rank=r, world=world)
                assert 0 <= b <= e <= spp
                seen += list(range(b, e))
            assert seen == list(range(spp))


def test_row_slices_partition_rows():
    for world in (1, 2, 3, 8):
        for h in (1, 18, 225, 1080, 2160):
            rows = []
            for r in range(world):
                y0, y1 = row_slice(h, r, world)
                rows += list(range(y0, y1))
            assert rows == list(range(h))


def test_interleaved_row_blocks_partition_rows():
    for world in (1, 2, 3, 8):
        for h in (1, 18, 225, 1080, 2160):
            for block in (4, 8, 32):
                rows = []
                for r in range(world):
                    blocks = row_blocks(h, r, world, block)
                    rows += [y for y0, y1 in blocks for y in range(y0, y1)]
                    # the one-call window maps onto the same rows: local row ly -> y0 + (ly // block) * stride + ly % block
                    y0, y1, stride = interleaved_window(h, r, world, block)
                    assert (y1 - y0) % block == 0 and stride == world * block
                    mapped = [y0 + (ly // block) * stride + ly % block for ly in range(y1 - y0)]
                    assert [y for y in mapped if y < h] == [y for a, b in blocks for y in range(a, b)]
                assert sorted(rows) == list(range(h))


_WORKER = r'''
import os, sys
sys.path.insert(0, %(repo)r); sys.path.insert(0, os.path.join(%(repo)r, "oracle")); sys.path.insert(0, os.path.join(%(repo)r, "tests"))
import numpy as np, torch, torch.distributed as dist
from path_tracer_python_b200 import scenes
import golden_fixtures
from path_tracer_python_b200.distributed import sample_slice, reduce_accum, row_blocks, gather_rows
import oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cs, meta, out = golden_fixtures.load_fixture("wavefront_comparison")
S = O.OracleScene(cs); cam = O.make_camera(meta["camera"], meta["background"], meta["max_depth"])
W, H, spp, seed = meta["width"], meta["height"], meta["spp"], meta["rng_seed"]
b, e = sample_slice(spp)
# the oracle stands in for the GPU kernels here: this test is about the split + reduce plumbing
part, _ = O.render(S, cam, W, H, e - b, seed=seed, sample_begin=b, nthreads=1)
acc = torch.from_numpy(part.copy())
reduce_accum(acc)
ref = out["accum_megakernel"]
err = float(np.abs(acc.numpy() - ref).max())
ok = np.allclose(acc.numpy(), ref, rtol=1e-5, atol=1e-5)
# tile split: every rank renders all samples of its interleaved row blocks; gather_rows assembles the
# frame without adding anything -> bit-identical to the single-process result
for dst in (None, 0):
    tile = np.zeros((H, W, 3), np.float32)
    for y0, y1 in row_blocks(H, rank, world, 4):
        O.render(S, cam, W, H, spp, seed=seed, accum=tile, y0=y0, y1=y1, nthreads=1)
    tt = torch.from_numpy(tile)
    gather_rows(tt, H, block=4, dst=dst)
    if dst is None or rank == dst:
        ok = ok and np.array_equal(tt.numpy(), ref)
print(f"RANK{rank} slice=({b},{e}) maxerr={err:.3e} ok={ok}")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_gloo_world2_spp_split_reduces_to_single_gpu_result(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"repo": REPO})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "RANK0" in r.stdout and "RANK1" in r.stdout and "ok=True" in r.stdout


def test_auto_samples_in_flight_sizes_batches_by_paths():
    """Engine batch sizing (host logic): ~66 M paths per batch, equal batches, library limit 256."""
    from path_tracer_python_b200.engine import auto_samples_in_flight as f
    assert f(1920, 1080, 256) == 32          # BASELINE config 2: 8 batches of 32
    assert f(1920, 1080, 16) == 16           # fewer samples than one batch holds
    assert f(1024, 1024, 64) == 32           # would fit one batch: split in two (tail / primary overlap)
    assert f(3840, 2160, 32) == 8
    assert f(400, 225, 16) == 16
    assert f(64, 64, 100000) == 256          # tiny window: the library's limit
    assert f(1920, 1080, 100) == 34          # 3 equal batches (34 + 34 + 32), not 32 + 32 + 32 + 4
    assert f(8192, 8192, 4) == 1
    # one rank's share of config 2 on 8 GPUs (32 samples): two batches of 16, so that the first batch's tail
    # overlaps the second batch's primary wave; small jobs are not split
    assert f(1920, 1080, 32) == 16 and f(1920, 1080, 24) == 24 and f(1920, 1080, 30) == 15
    assert f(1920, 1080, 32, min_batches=4) == 8 and f(400, 225, 16, min_batches=2) == 8
