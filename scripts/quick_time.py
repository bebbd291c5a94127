"""Ad-hoc device timings of the main ops (development aid, not the bench contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch3d_pointops_b200 import _C
from pytorch3d_pointops_b200.functions import ball_query, sample_farthest_points, knn_points
from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance

dev = "cuda:0"

def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]

g = torch.Generator().manual_seed(0)
which = sys.argv[1:] or ["knn", "chamfer", "fps", "bq"]
if "knn" in which:
    for (N, P, K) in [(32, 16384, 16), (32, 16384, 1), (32, 16384, 32), (8, 8192, 16)]:
        p = torch.rand(N, P, 3, generator=g).to(dev)
        L = torch.full((N,), P, device=dev)
        mn, med = timeit(lambda: _C.knn_points_idx(p, p, L, L, 2, K, -1))
        import ctypes
        from pytorch3d_pointops_b200 import _lib
        lib = _lib.load(); lib.pops_profile_reset(); lib.pops_profile_enable(1)
        for _ in range(3): _C.knn_points_idx(p, p, L, L, 2, K, -1)
        torch.cuda.synchronize()
        nl, ms = ctypes.c_int64(0), ctypes.c_double(0)
        lib.pops_profile_read(b"knn_scan", ctypes.byref(nl), ctypes.byref(ms)); lib.pops_profile_enable(0)
        print(f"   scan kernel alone: {ms.value / max(1, nl.value):.3f} ms")
        pairs = N * P * P
        print(f"knn N={N} P={P} K={K}: min {mn:.3f} ms med {med:.3f} ms  {N*P/mn/1e3:.1f} Mq/s  {pairs*9/mn/1e9:.1f} TFLOP/s-alg")
if "chamfer" in which:
    N, P = 32, 8192
    x = torch.rand(N, P, 3, generator=g).to(dev).requires_grad_(True)
    y = torch.rand(N, P, 3, generator=g).to(dev).requires_grad_(True)
    xl = torch.randint(4096, P + 1, (N,), generator=g).to(dev)
    yl = torch.randint(4096, P + 1, (N,), generator=g).to(dev)
    xn = torch.nn.functional.normalize(torch.randn(N, P, 3, generator=g), dim=-1).to(dev).requires_grad_(True)
    yn = torch.nn.functional.normalize(torch.randn(N, P, 3, generator=g), dim=-1).to(dev).requires_grad_(True)
    xc = torch.rand(N, P, 3, generator=g).to(dev).requires_grad_(True)
    yc = torch.rand(N, P, 3, generator=g).to(dev).requires_grad_(True)
    def step():
        loss, lf = chamfer_distance(x, y, x_lengths=xl, y_lengths=yl, x_features={"normals": xn, "colors": xc},
                                    y_features={"normals": yn, "colors": yc}, feature_names=["normals", "colors"])
        (loss + lf["normals"] + lf["colors"]).backward()
    mn, med = timeit(step)
    print(f"chamfer fwd+bwd N={N} P<={P}: min {mn:.3f} ms med {med:.3f}  {N/mn*1e3:.0f} pairs/s")
if "fps" in which:
    for (N, P, K) in [(64, 65536, 1024), (8, 65536, 1024), (64, 4096, 512)]:
        pts = torch.rand(N, P, 3, generator=g).to(dev)
        mn, med = timeit(lambda: sample_farthest_points(pts, K=K), n=3, warm=1)
        print(f"fps N={N} P={P} K={K}: min {mn:.3f} ms  {N*K/mn/1e3:.2f} Msamples/s  {mn*1e3/K:.2f} us/iter")
if "bq" in which:
    for (N, P1, P2) in [(128, 16384, 16384), (128, 4096, 16384)]:
        p2 = torch.rand(N, P2, 3, generator=g).to(dev)
        p1 = p2[:, :P1].contiguous()
        mn, med = timeit(lambda: ball_query(p1, p2, K=32, radius=0.1), n=3, warm=1)
        print(f"ball_query N={N} P1={P1} P2={P2} K=32 r=0.1 (+gather): min {mn:.3f} ms  {N*P1/mn/1e3:.1f} Mq/s")
