import sys; sys.path.insert(0, "/root/repo")
import torch
from pytorch3d_pointops_b200 import _C, _lib
lib = _lib.load(); dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
p = torch.rand(32, 16384, 3, generator=g).to(dev); L = torch.full((32,), 16384, dtype=torch.int64, device=dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for q in (0,):
    lib.pops_set_option(b"knn_q", q)
    for K in (1, 4, 6, 8, 12, 16, 32):
        print(f"Q={q} K={K}: {timeit(lambda: _C.knn_points_idx(p, p, L, L, 2, K, -1)):.4f} ms")
