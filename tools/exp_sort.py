"""Experiment: how much does the ORDER of incoherent bounce rays matter for traversal throughput?
Builds wave-1-like rays (cosine-weighted bounces off the primary hits of config 2) and times the
first-hit kernel on them in different orders (development aid; not bit-exact shading)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))   # golden_fixtures
import torch
from path_tracer_python_b200.engine import Engine
from path_tracer_python_b200 import scenes
import golden_fixtures
W = 1920
e = Engine(0)
cs, meta = golden_fixtures.load_config("vol2_final_scene", W, 16 / 9)
e.set_scene(cs); e.set_camera(meta["camera"], (0, 0, 0), 50)
H = meta["height"]
g = torch.Generator(device="cuda").manual_seed(1)
O, D = [], []
for s in range(int(os.environ.get("NS", 4))):
    o, d = e.generate_camera_rays(W, H, 7, s)
    t, prim, p, n = e.first_hit(o, d)
    hit = prim >= 0
    # flip normal against the ray, cosine-weighted direction
    n = torch.where(((n * d).sum(1, keepdim=True) > 0), -n, n)
    r1 = torch.rand(len(n), device="cuda", generator=g); r2 = torch.rand(len(n), device="cuda", generator=g)
    phi = 2 * torch.pi * r1; rr = r2.sqrt()
    a = torch.where(n[:, :1].abs() > 0.9, torch.tensor([0., 1., 0.], device="cuda"), torch.tensor([1., 0., 0.], device="cuda")).expand_as(n)
    tt = torch.linalg.cross(n, a); tt = tt / tt.norm(dim=1, keepdim=True); bb = torch.linalg.cross(n, tt)
    nd = tt * (rr * phi.cos())[:, None] + bb * (rr * phi.sin())[:, None] + n * (1 - r2).sqrt()[:, None]
    O.append(p[hit]); D.append(nd[hit])
O = torch.cat(O); D = torch.cat(D)
N = len(O)
print("bounce rays:", N)
def timeit(o, d, reps=3):
    e.first_hit(o, d, want_pn=False); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); e.first_hit(o, d, want_pn=False); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return N / best / 1e6
def morton(o, bits=10):
    lo, hi = o.min(0).values, o.max(0).values
    q = ((o - lo) / (hi - lo + 1e-9) * (2 ** bits - 1)).long()
    code = torch.zeros(len(o), dtype=torch.long, device="cuda")
    for b in range(bits):
        for ax in range(3):
            code |= ((q[:, ax] >> b) & 1) << (3 * b + ax)
    return code
octant = ((D[:, 0] < 0).long() | ((D[:, 1] < 0).long() << 1) | ((D[:, 2] < 0).long() << 2))
mc = morton(O)
# finer direction key: quantised direction (6 x 8 x 8 cube-map cells)
ax = D.abs().argmax(1); major = torch.gather(D, 1, ax[:, None]).squeeze(1)
face = ax * 2 + (major < 0).long()
uv = torch.stack([D[:, 1] / D.abs().max(1).values, D[:, 2] / D.abs().max(1).values], 1)
cell = face * 16 + ((uv[:, 0] * 0.5 + 0.5) * 3.999).long() * 4 + ((uv[:, 1] * 0.5 + 0.5) * 3.999).long()
def dircell(k):
    return face * k * k + ((uv[:, 0] * 0.5 + 0.5) * (k - 1e-3)).long() * k + ((uv[:, 1] * 0.5 + 0.5) * (k - 1e-3)).long()
tile = torch.arange(N, device="cuda") // 2048          # what one CTA of a binning pass would see
orders = {
    "faces(6) stable": torch.sort(face, stable=True).indices,
    "cells(24) stable": torch.sort(dircell(2), stable=True).indices,
    "cells(96) stable": torch.sort(dircell(4), stable=True).indices,
    "cells(384) stable": torch.sort(dircell(8), stable=True).indices,
    "cells(1536) stable": torch.sort(dircell(16), stable=True).indices,
    "cells(96) + morton 3 bits/axis": torch.argsort(dircell(4) * 512 + (mc >> 21), stable=True),
    "cells(96) + morton 5 bits/axis": torch.argsort(dircell(4) * 32768 + (mc >> 15), stable=True),
    "morton 4 bits/axis + cells(96)": torch.argsort((mc >> 18) * 96 + dircell(4), stable=True),
    "morton 6 bits/axis + cells(24)": torch.argsort((mc >> 12) * 24 + dircell(2), stable=True),
    "as generated (pixel order)": torch.arange(N, device="cuda"),
    "random permutation": torch.randperm(N, device="cuda", generator=g),
    "octant only (stable)": torch.sort(octant, stable=True).indices,
    "morton(origin) only": torch.argsort(mc),
    "octant, then morton": torch.argsort(octant * (1 << 32) + mc),
    "morton(5 bits/axis), then octant": torch.argsort(((mc >> 15) << 3) + octant, stable=True),
    "dir cell(96), then morton": torch.argsort(cell * (1 << 32) + mc),
}
for name, idx in orders.items():
    print(f"{name:36s} {timeit(O[idx].contiguous(), D[idx].contiguous()):8.1f} Mrays/s", flush=True)
