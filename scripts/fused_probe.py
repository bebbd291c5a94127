import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from pytorch3d_pointops_b200 import _C, _lib
from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance
lib = _lib.load(); dev = torch.device("cuda:0")
ch = {k: v.to(dev) for k, v in bench.make_chamfer_inputs(0).items()}
for k in ("x", "y", "xn", "yn", "xc", "yc"):
    ch[k].requires_grad_(True)
g = torch.Generator().manual_seed(0)
p = torch.rand(32, 16384, 3, generator=g).to(dev); L = torch.full((32,), 16384, dtype=torch.int64, device=dev)
p8 = torch.rand(64, 8192, 3, generator=g).to(dev); L8 = torch.full((64,), 8192, dtype=torch.int64, device=dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def step():
    for k in ("x", "y", "xn", "yn", "xc", "yc"):
        ch[k].grad = None
    loss, lf = chamfer_distance(ch["x"], ch["y"], x_lengths=ch["xl"], y_lengths=ch["yl"],
                                x_features={"normals": ch["xn"], "colors": ch["xc"]},
                                y_features={"normals": ch["yn"], "colors": ch["yc"]}, feature_names=["normals", "colors"])
    (loss + lf["normals"] + lf["colors"]).backward()
def wall(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
with torch.no_grad():
    xd, yd = ch['x'].detach(), ch['y'].detach()
for f, items in ((0, 8), (1, 8), (1, 16), (0, 8), (1, 8), (1, 16)):
    lib.pops_set_option(b"knn_fused_prepass", f)
    lib.pops_set_option(b"knn_fused_items", items)
    tp = timeit(lambda: _C.knn_points_idx_pair(xd, yd, ch['xl'], ch['yl'], 2, 1))
    t16 = timeit(lambda: _C.knn_points_idx(p, p, L, L, 2, 16, -1))
    t8 = timeit(lambda: _C.knn_points_idx(p8, p8, L8, L8, 2, 16, -1))
    print(f"fused={f} items={items}: chamfer pair {tp:.4f} ms  T shape K=16 {t16:.4f} ms  64x8192 K=16 {t8:.4f} ms  chamfer step wall {wall(step):.4f} ms", flush=True)
