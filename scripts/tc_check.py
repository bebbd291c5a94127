"""Development aid: tensor-core KNN path vs the CPU oracle on a few shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as O
from pytorch3d_pointops_b200 import _C

def check(N, P1, P2, D, K, ragged=False, seed=0, scale=1.0, kind="randn"):
    g = torch.Generator().manual_seed(seed)
    if kind == "randn":
        a = torch.randn(N, P1, D, generator=g) * scale
        b = torch.randn(N, P2, D, generator=g) * scale
    else:  # integer grid: massive ties
        a = torch.randint(0, 3, (N, P1, D), generator=g).float()
        b = torch.randint(0, 3, (N, P2, D), generator=g).float()
    l1 = torch.full((N,), P1, dtype=torch.int64)
    l2 = torch.full((N,), P2, dtype=torch.int64)
    if ragged:
        l1 = torch.randint(1, P1 + 1, (N,), generator=g)
        l2 = torch.randint(1, P2 + 1, (N,), generator=g)
        l2[0] = min(P2, 7)
    oi, od = O.knn_points_idx(a, b, l1, l2, 2, K)
    idx, d = _C.knn_points_idx(a.cuda(), b.cuda(), l1.cuda(), l2.cuda(), 2, K, -1)
    torch.cuda.synchronize()
    ok_i = torch.equal(idx.cpu(), oi)
    ok_d = torch.equal(d.cpu(), od)
    print(f"N={N} P1={P1} P2={P2} D={D} K={K} ragged={ragged} kind={kind}: idx {'OK' if ok_i else 'MISMATCH'} dists {'OK' if ok_d else 'MISMATCH'}", flush=True)
    if not ok_i:
        bad = (idx.cpu() != oi).nonzero()
        print("  first mismatches:", bad[:5].tolist(), idx.cpu()[tuple(bad[0][:2])].tolist(), oi[tuple(bad[0][:2])].tolist())
    return ok_i and ok_d

ok = True
ok &= check(1, 128, 512, 128, 16)
ok &= check(2, 300, 1000, 128, 16)
ok &= check(2, 300, 1000, 64, 8, ragged=True)
ok &= check(3, 257, 2049, 32, 1)
ok &= check(2, 200, 1500, 100, 5, ragged=True)
ok &= check(2, 200, 1500, 256, 16)
ok &= check(1, 130, 700, 128, 16, kind="grid")
ok &= check(2, 1000, 5000, 128, 16, scale=100.0)
print("ALL OK" if ok else "FAILED")
if ok and len(sys.argv) > 1:
    N, P, D, K = 16, 32768, 128, 16
    g = torch.Generator().manual_seed(4)
    x = torch.randn(N, P, D, generator=g).cuda()
    L = torch.full((N,), P, device="cuda")
    for _ in range(2):
        _C.knn_points_idx(x, x, L, L, 2, K, -1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); _C.knn_points_idx(x, x, L, L, 2, K, -1); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"C5 shape: {ms:.2f} ms  {N*P/ms/1e3:.1f} Mq/s  GEMM-equivalent {2*D*N*P*P/ms/1e9:.0f} TFLOP/s")
