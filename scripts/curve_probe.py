"""Morton vs Hilbert ordering in the pre-pass: blocks visited and time, chamfer K=1 shape and T shape."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from pytorch3d_pointops_b200 import _C, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
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
def stats(fn):
    out = (ctypes.c_ulonglong * 8)()
    lib.pops_set_option(b"knn_stats", 1); lib.pops_knn_debug_stats(out); fn(); lib.pops_knn_debug_stats(out)
    lib.pops_set_option(b"knn_stats", 0)
    f, s, fl, cg, ne, w = [int(x) for x in out[:6]]
    return f"fetched/warp {f/w:.1f} scanned/warp {s/w:.1f} flush rounds/warp {fl/w:.1f}"
res = {}
for curve in (0, 1):
    lib.pops_set_option(b"knn_curve", curve)
    f1 = lambda: _C.knn_points_idx_pair(ch['x'], ch['y'], ch['xl'], ch['yl'], 2, 1)
    f16 = lambda: _C.knn_points_idx(p, p, L, L, 2, 16, -1)
    f4 = lambda: _C.knn_points_idx(p, p, L, L, 2, 4, -1)
    f32 = lambda: _C.knn_points_idx(p, p, L, L, 2, 32, -1)
    res[curve] = (f1(), f16())
    print(f"curve={curve} chamfer pair K=1: {timeit(f1):.4f} ms  [{stats(f1)}]")
    print(f"curve={curve} T shape K=16:     {timeit(f16):.4f} ms  [{stats(f16)}]")
    print(f"curve={curve} T shape K=4:      {timeit(f4):.4f} ms   K=32: {timeit(f32):.4f} ms")
same = all(torch.equal(a, b) for a, b in zip(res[0][0] + res[0][1], res[1][0] + res[1][1]))
print("results identical:", same)
