import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from pytorch3d_pointops_b200 import _C, _lib
lib = _lib.load(); dev = torch.device("cuda:0")
ch = {k: v.to(dev) for k, v in bench.make_chamfer_inputs(0).items()}
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
ref = None
for q in (4, 2, 1):
    lib.pops_set_option(b"knn_q", q)
    out = _C.knn_points_idx_pair(ch['x'], ch['y'], ch['xl'], ch['yl'], 2, 1)
    if ref is None: ref = out
    same = all(torch.equal(a, b) for a, b in zip(ref, out))
    tp = timeit(lambda: _C.knn_points_idx_pair(ch['x'], ch['y'], ch['xl'], ch['yl'], 2, 1))
    ts = [timeit(lambda: _C.knn_points_idx(p, p, L, L, 2, K, -1)) for K in (1, 4, 16, 32)]
    print(f"Q={q}: chamfer pair {tp:.4f} ms (same={same});  T shape K=1/4/16/32: " + " ".join(f"{t:.3f}" for t in ts))
