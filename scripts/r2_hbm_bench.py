"""Development probe (round 2): CUDA-event timing of the HBM-bound kernels on the north_star shapes,
with the SURVEY 8(d) byte formulas.  bench.py carries the same measurements in its JSON line; this
script exists to be run under ncu on a short command line.

    python scripts/r2_hbm_bench.py [reps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_pointops_b200 import _C  # noqa: E402

DEV = torch.device("cuda", 0)
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 10
PEAK = 6556.5
flush = torch.empty(384 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, reps=REPS):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2]


def report(name, ms, alg_bytes, dram_bytes):
    print(f"{name:34s} {ms * 1e3:9.1f} us  algorithmic {alg_bytes / ms / 1e6:8.1f} GB/s ({alg_bytes / ms / 1e6 / PEAK:.2f})"
          f"  compulsory-DRAM {dram_bytes / ms / 1e6:8.1f} GB/s ({dram_bytes / ms / 1e6 / PEAK:.2f})", flush=True)


def main():
    g = torch.Generator().manual_seed(3)
    # ---- gather: C4 (B=128, P=16384, K=32, U=3), ball-query indices (masked) and KNN indices ----------
    N, P, K = 128, 16384, 32
    x = torch.rand(N, P, 3, generator=g).to(DEV)
    idx = torch.randint(0, P, (N, P, K), generator=g).to(DEV)
    rows = N * P * K
    for mode, nm in ((_C.GATHER_MASKED, "masked"), (_C.GATHER_KNN, "knn")):
        ms = timed(lambda: _C.gather(x, idx, None, mode))
        report(f"gather {nm} C4 U=3", ms, rows * (8 + 12 + 12), rows * (8 + 12) + x.numel() * 4)
    x3 = torch.rand(32, 16384, 3, generator=g).to(DEV)
    L3 = torch.full((32,), 16384, dtype=torch.int64, device=DEV)
    idx3, _ = _C.knn_points_idx(x3, x3, L3, L3, 2, 16, -1)
    r3 = 32 * 16384 * 16
    ms = timed(lambda: _C.gather(x3, idx3, L3, _C.GATHER_KNN))
    report("gather knn T U=3 (KNN indices)", ms, r3 * (8 + 12 + 12), r3 * (8 + 12) + x3.numel() * 4)
    x16 = torch.rand(32, 16384, 16, generator=g).to(DEV)
    idx16 = torch.randint(0, 16384, (32, 16384, 16), generator=g).to(DEV)
    r16 = 32 * 16384 * 16
    ms = timed(lambda: _C.gather(x16, idx16, None, _C.GATHER_KNN))
    report("gather knn T U=16", ms, r16 * (8 + 64 + 64), r16 * (8 + 64) + x16.numel() * 4)
    del x, idx, x16, idx16
    # ---- packed <-> padded: 32 ragged clouds of <= 65536 points, D=3 and D=16 -----------------------
    for D in (3, 16):
        lens = torch.randint(32768, 65537, (64,), generator=g)
        first = (torch.cumsum(lens, 0) - lens).to(DEV)
        F, mx = int(lens.sum()), int(lens.max())
        packed = torch.rand(F, D, generator=g).to(DEV)
        ms = timed(lambda: _C.packed_to_padded(packed, first, mx))
        by = F * D * 4 + 64 * mx * D * 4
        report(f"packed_to_padded 64x<=65536 D={D}", ms, by, by)
        padded = _C.packed_to_padded(packed, first, mx)
        ms = timed(lambda: _C.padded_to_packed(padded, first, F))
        report(f"padded_to_packed 64x<=65536 D={D}", ms, 2 * F * D * 4, 2 * F * D * 4)
        del packed, padded
    # ---- knn backward: T shape (B=32, P=16384, K=16, D=3) ---------------------------------------------
    N, P, K = 32, 16384, 16
    p = torch.rand(N, P, 3, generator=g).to(DEV)
    L = torch.full((N,), P, dtype=torch.int64, device=DEV)
    idx, _ = _C.knn_points_idx(p, p, L, L, 2, K, -1)
    gd = torch.rand(N, P, K, generator=g).to(DEV)
    ms = timed(lambda: _C.knn_points_backward(p, p, L, L, idx, 2, gd))
    e = N * P * K
    alg = e * 12 + 2 * e * 3 * 4 + 2 * N * P * 3 * 4
    dram = e * 12 + 3 * N * P * 3 * 4
    report("knn_backward T (D=3, K=16)", ms, alg, dram)
    from pytorch3d_pointops_b200 import _lib
    lib = _lib.load()
    lib.pops_set_option(b"knn_backward_rows", 0)
    ms = timed(lambda: _C.knn_points_backward(p, p, L, L, idx, 2, gd))
    report("knn_backward T, r1 kernel", ms, alg, dram)
    lib.pops_set_option(b"knn_backward_rows", 1)


if __name__ == "__main__":
    main()
