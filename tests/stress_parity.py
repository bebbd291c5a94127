"""One-off stress run: random shapes, lengths, K and geometries through knn_points_idx (D = 3, L2: the pruned
search, both pre-pass forms), the two-sided pair search and ball_query, each compared bit for bit with the
CPU oracle.  Test infrastructure like the rest of tests/ (imports oracle/); run by hand on a GPU box:

    python tests/stress_parity.py [seed] [cases]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as O
from pytorch3d_pointops_b200 import _C, _lib

O.build()
lib = _lib.load()
dev = "cuda:0"
gen = torch.Generator().manual_seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 60
bad = 0
for case in range(ncase):
    N = int(torch.randint(1, 5, (1,), generator=gen))
    P1 = int(torch.randint(1, 3000, (1,), generator=gen))
    P2 = int(torch.randint(64, 6000, (1,), generator=gen))
    K = [1, 2, 4, 5, 8, 11, 16, 17, 32][int(torch.randint(0, 9, (1,), generator=gen))]
    geo = int(torch.randint(0, 4, (1,), generator=gen))
    if geo == 0:
        p1, p2 = torch.rand(N, P1, 3, generator=gen), torch.rand(N, P2, 3, generator=gen)
    elif geo == 1:  # integer grid: masses of exact ties
        p1 = torch.randint(0, 6, (N, P1, 3), generator=gen).float()
        p2 = torch.randint(0, 6, (N, P2, 3), generator=gen).float()
    elif geo == 2:  # clusters + offset
        c = 100.0 * torch.randn(N, 5, 3, generator=gen)
        p1 = c[:, torch.randint(0, 5, (P1,), generator=gen)] + 0.1 * torch.randn(N, P1, 3, generator=gen) + 1e3
        p2 = c[:, torch.randint(0, 5, (P2,), generator=gen)] + 0.1 * torch.randn(N, P2, 3, generator=gen) + 1e3
    else:  # thin slab
        p1, p2 = torch.randn(N, P1, 3, generator=gen), torch.randn(N, P2, 3, generator=gen)
        p1[..., 2] *= 1e-4
        p2[..., 2] *= 1e-4
    l1 = torch.randint(0, P1 + 1, (N,), generator=gen)
    l2 = torch.randint(0, P2 + 1, (N,), generator=gen)
    l1[0], l2[0] = P1, P2
    oi, od = O.knn_points_idx(p1, p2, l1, l2, 2, K)
    a = [t.to(dev) for t in (p1, p2, l1, l2)]
    for fused in (1, 0):
        lib.pops_set_option(b"knn_fused_prepass", fused)
        gi, gd = _C.knn_points_idx(a[0], a[1], a[2], a[3], 2, K, -1)
        ok = torch.equal(gi.cpu(), oi) and torch.equal(gd.cpu(), od)
        if not ok:
            bad += 1
            print(f"MISMATCH knn case {case} fused={fused}: N={N} P1={P1} P2={P2} K={K} geo={geo}", flush=True)
    lib.pops_set_option(b"knn_fused_prepass", 1)
    if K == 1 and P1 >= 64:
        i12, d12, i21, d21 = _C.knn_points_idx_pair(a[0], a[1], a[2], a[3], 2, 1)
        oi2, od2 = O.knn_points_idx(p2, p1, l2, l1, 2, 1)
        if not (torch.equal(i12.cpu(), oi) and torch.equal(d12.cpu(), od) and torch.equal(i21.cpu(), oi2) and torch.equal(d21.cpu(), od2)):
            bad += 1
            print(f"MISMATCH pair case {case}: N={N} P1={P1} P2={P2} geo={geo}", flush=True)
    if case % 3 == 0 and K >= 4:
        r = {0: 0.08, 1: 1.5, 2: 0.3, 3: 0.2}[geo]
        bi, bd = O.ball_query_idx(p1, p2, l1, l2, K, r)
        for mode in (1, 0):
            lib.pops_set_option(b"bq_spatial", mode)
            gi, gd = _C.ball_query(a[0], a[1], a[2], a[3], K, r)
            if not (torch.equal(gi.cpu(), bi) and torch.equal(gd.cpu(), bd)):
                bad += 1
                print(f"MISMATCH ball_query case {case} mode={mode}: N={N} P1={P1} P2={P2} K={K} geo={geo} r={r}", flush=True)
        lib.pops_set_option(b"bq_spatial", -1)
print(f"{ncase} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
