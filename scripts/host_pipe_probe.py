import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch3d_pointops_b200 import _C
from pytorch3d_pointops_b200.host import HostKnn
dev = torch.device("cuda:0")
B, P, K = 32, 16384, 16
g = torch.Generator().manual_seed(0)
p = torch.rand(B, P, 3, generator=g).pin_memory()
L = torch.full((B,), P, dtype=torch.int64).pin_memory()
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
pd, Ld = p.to(dev), L.to(dev)
idx, d = _C.knn_points_idx(pd, pd, Ld, Ld, 2, K, -1)
oi = torch.empty(idx.shape, dtype=idx.dtype).pin_memory(); od = torch.empty(d.shape, dtype=d.dtype).pin_memory()
print("compute all   ", timeit(lambda: _C.knn_points_idx(pd, pd, Ld, Ld, 2, K, -1)))
print("compute 8 of32", timeit(lambda: _C.knn_points_idx(pd[:8], pd[:8], Ld[:8], Ld[:8], 2, K, -1)))
print("d2h 100MB     ", timeit(lambda: (oi.copy_(idx, non_blocking=True), od.copy_(d, non_blocking=True))))
print("h2d 6MB       ", timeit(lambda: p.to(dev, non_blocking=True)))
for s in (2, 4, 6, 8, 10, [3, 5, 6, 6, 6, 6], [2, 6, 6, 6, 6, 6], [4, 7, 7, 7, 7]):
    hk = HostKnn(B, P, P, 3, K, dev, slices=s)
    print(f"pipeline s={s}  ", timeit(lambda: hk(p, None, L)))
