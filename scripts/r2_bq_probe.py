"""Development probe (round 2): Hilbert-ordered ball query (bq_spatial=1) vs the index-order scan
(bq_spatial=0): same rows, time per call over radius / K / cloud shape."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_pointops_b200 import _C, _lib  # noqa: E402

DEV = torch.device("cuda", 0)
lib = _lib.load()
flush = torch.empty(384 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, reps=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return out, sorted(ts)[len(ts) // 2]


g = torch.Generator().manual_seed(0)
N = int(os.environ.get("BQ_N", "32"))
shapes = [("uniform", torch.rand(N, 16384, 3, generator=g))]
s = torch.randn(N, 16384, 3, generator=g)
shapes.append(("sphere", s / s.norm(dim=-1, keepdim=True) * 0.5 + 0.5))
for name, pts in shapes:
    p = pts.to(DEV)
    L = torch.full((N,), p.shape[1], dtype=torch.int64, device=DEV)
    for K in (32, 64, 16):
        for r in (0.05, 0.1, 0.15, 0.2, 0.3):
            res = {}
            for mode in (0, 1, -1):
                lib.pops_set_option(b"bq_spatial", mode)
                res[mode] = timed(lambda: _C.ball_query(p, p, L, L, K, r))
            eq = all(torch.equal(x, y) for x, y in zip(res[0][0], res[1][0])) and all(
                torch.equal(x, y) for x, y in zip(res[0][0], res[-1][0]))
            hits = (res[0][0][0] >= 0).sum(-1).float().mean().item()
            print(f"{name} N={N} K={K} r={r}: scan {res[0][1]:7.3f} ms   spatial {res[1][1]:7.3f} ms   auto {res[-1][1]:7.3f} ms"
                  f"   filled {hits:5.1f}  equal={eq}", flush=True)
lib.pops_set_option(b"bq_spatial", -1)
