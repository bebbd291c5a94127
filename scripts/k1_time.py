"""K=1 pair search timing on the chamfer shape (kernel-level, via the profile hooks)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from pytorch3d_pointops_b200 import _C, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
ch = {k: v.to(dev) for k, v in bench.make_chamfer_inputs(0).items()}
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for q in (4, 2):
    lib.pops_set_option(b"knn_q", q)
    print(f"Q={q} pair   {timeit(lambda: _C.knn_points_idx_pair(ch['x'], ch['y'], ch['xl'], ch['yl'], 2, 1)):.4f} ms")
    print(f"Q={q} single {timeit(lambda: _C.knn_points_idx(ch['x'], ch['y'], ch['xl'], ch['yl'], 2, 1, -1)):.4f} ms")
lib.pops_set_option(b"knn_q", 4)
ks = _C.KnnSliced(ch['x'], ch['y'], ch['xl'], ch['yl'], 2, 1)
ks.prepare()
print(f"search only {timeit(lambda: ks.search(0, 32)):.4f} ms   prepare only {timeit(ks.prepare):.4f} ms")
