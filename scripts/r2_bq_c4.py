"""Development probe: ball query on the C4 shape (B=128, P=16384, K=32, r=0.1), scan vs auto."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_pointops_b200 import _C, _lib  # noqa: E402

lib = _lib.load()
g = torch.Generator().manual_seed(0)
N = int(os.environ.get("BQ_N", "128"))
p = torch.rand(N, 16384, 3, generator=g).cuda()
L = torch.full((N,), 16384, dtype=torch.int64, device="cuda")
for mode in ([int(os.environ["BQ_MODE"])] if "BQ_MODE" in os.environ else [0, -1]):
    lib.pops_set_option(b"bq_spatial", mode)
    for _ in range(2):
        out = _C.ball_query(p, p, L, L, 32, 0.1)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = _C.ball_query(p, p, L, L, 32, 0.1); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"N={N} bq_spatial={mode}: {sorted(ts)[2]:.3f} ms", flush=True)
lib.pops_set_option(b"bq_spatial", -1)
