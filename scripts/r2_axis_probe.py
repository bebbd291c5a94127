"""Development probe (round 2): T-shape search kernel / step time against the curve-code resolution
(option knn_axis_bits) -- how coarse can the ordering grid be before the block boxes loosen."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_pointops_b200 import _C, _lib  # noqa: E402

DEV = torch.device("cuda", 0)
lib = _lib.load()
flush = torch.empty(384 << 20, dtype=torch.uint8, device=DEV)


def kernel_ms(name=b"knn_scan"):
    n_, ms_ = ctypes.c_int64(0), ctypes.c_double(0.0)
    lib.pops_profile_read(name, ctypes.byref(n_), ctypes.byref(ms_))
    return ms_.value / max(1, n_.value)


def run(p, L, K, reps=10):
    for _ in range(2):
        out = _C.knn_points_idx(p, p, L, L, 2, K, -1)
    torch.cuda.synchronize()
    lib.pops_profile_reset()
    lib.pops_profile_enable(1)
    evs = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = _C.knn_points_idx(p, p, L, L, 2, K, -1)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    lib.pops_profile_enable(0)
    ms = kernel_ms()
    step = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
    lib.pops_profile_reset()
    return out, ms, step


g = torch.Generator().manual_seed(0)
p = torch.rand(32, 16384, 3, generator=g).to(DEV)
L = torch.full((32,), 16384, dtype=torch.int64, device=DEV)
# a surface-like cloud: points on a sphere shell (2-D manifold), the usual shape of real data
s = torch.randn(32, 16384, 3, generator=g)
s = (s / s.norm(dim=-1, keepdim=True)).to(DEV)
for name, pts in (("uniform", p), ("sphere", s)):
    ref = None
    for bits in (6, 5, 4, 3):
        lib.pops_set_option(b"knn_axis_bits", bits)
        for K in (16, 1):
            (i, d), ms, st = run(pts, L, K)
            if bits == 6:
                ref = ref or {}
                ref[K] = (i, d)
            eq = torch.equal(i, ref[K][0]) and torch.equal(d, ref[K][1])
            print(f"{name} axis_bits={bits} K={K:2d}: kernel {ms * 1e3:7.1f} us  step {st * 1e3:7.1f} us  equal={eq}")
lib.pops_set_option(b"knn_axis_bits", 0)
