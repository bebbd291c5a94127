"""Q (queries per thread) across shapes: make sure the default is not a regression anywhere."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch3d_pointops_b200 import _C, _lib
lib = _lib.load(); dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
def timeit(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
shapes = [(1, 300000, 300000, 16, "uniform"), (4, 100000, 100000, 8, "uniform"), (64, 2048, 2048, 16, "uniform"),
          (32, 1024, 16384, 16, "uniform"), (32, 16384, 1024, 4, "uniform"), (8, 65536, 65536, 1, "uniform"),
          (32, 16384, 16384, 16, "surface"), (32, 16384, 16384, 16, "clustered"), (128, 4096, 4096, 16, "uniform")]
for (N, P1, P2, K, kind) in shapes:
    p2 = torch.rand(N, P2, 3, generator=g)
    if kind == "surface":
        p2[..., 2] = 0.3 * torch.sin(6 * p2[..., 0]) * torch.cos(6 * p2[..., 1])
    if kind == "clustered":
        c = torch.rand(N, 16, 3, generator=g)
        p2 = c[:, torch.randint(0, 16, (P2,), generator=g)] + 0.02 * torch.randn(N, P2, 3, generator=g)
    p2 = p2.to(dev)
    p1 = p2 if P1 == P2 else torch.rand(N, P1, 3, generator=g).to(dev)
    l1 = torch.full((N,), P1, dtype=torch.int64, device=dev); l2 = l1 if P1 == P2 else torch.full((N,), P2, dtype=torch.int64, device=dev)
    ts = []
    for q in (1, 2, 4):
        lib.pops_set_option(b"knn_q", q)
        ts.append(timeit(lambda: _C.knn_points_idx(p1, p2, l1, l2, 2, K, -1)))
    print(f"N={N} P1={P1} P2={P2} K={K} {kind}: Q=1 {ts[0]:.3f}  Q=2 {ts[1]:.3f}  Q=4 {ts[2]:.3f} ms", flush=True)
lib.pops_set_option(b"knn_q", 0)
