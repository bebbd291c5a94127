import sys, os, ctypes
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
for bits in (4, 5, 6, 7, 8):
    lib.pops_set_option(b"knn_axis_bits", bits)
    t16 = timeit(lambda: _C.knn_points_idx(p, p, L, L, 2, 16, -1))
    t1 = timeit(lambda: _C.knn_points_idx_pair(ch['x'], ch['y'], ch['xl'], ch['yl'], 2, 1))
    print(f"axis_bits={bits}: T shape K=16 {t16:.4f} ms   chamfer pair {t1:.4f} ms")
