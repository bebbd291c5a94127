"""Development probe (round 2): cluster pre-pass (option knn_fused_prepass) against the multi-launch /
single-CTA pre-passes -- same results, step time per shape."""
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


def timed(fn, reps=10):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    lib.pops_profile_reset()
    lib.pops_profile_enable(1)
    evs = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    lib.pops_profile_enable(0)
    ms = (kernel_ms(), kernel_ms(b"knn_order"))
    step = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
    lib.pops_profile_reset()
    return out, ms, step


def same(a, b):
    return all(torch.equal(x, y) for x, y in zip(a, b))


g = torch.Generator().manual_seed(0)
cases = []
for name, N, P1, P2, K, ragged, selfk in (
    ("T uniform", 32, 16384, 16384, 16, False, True),
    ("T ragged", 32, 16384, 16384, 16, True, True),
    ("C2 pair-like", 32, 8192, 8192, 1, True, False),
    ("P1!=P2", 8, 5000, 12000, 8, True, False),
    ("one cloud 16k", 1, 16384, 16384, 16, False, True),
    ("one cloud 4k", 1, 4096, 4096, 8, False, True),
    ("small 500", 32, 500, 500, 8, False, True),
    ("64k self", 4, 65536, 65536, 16, True, True),
    ("30k two", 4, 30000, 32768, 4, True, False),
    ("128 clouds", 128, 16384, 16384, 16, False, True),
):
    p2 = torch.rand(N, P2, 3, generator=g).to(DEV)
    p1 = p2 if selfk else torch.rand(N, P1, 3, generator=g).to(DEV)
    if ragged:
        L2 = torch.randint(P2 // 2, P2 + 1, (N,), generator=g).to(DEV)
        L1 = L2 if selfk else torch.randint(P1 // 2, P1 + 1, (N,), generator=g).to(DEV)
    else:
        L2 = torch.full((N,), P2, dtype=torch.int64, device=DEV)
        L1 = L2 if selfk else torch.full((N,), P1, dtype=torch.int64, device=DEV)
    res = {}
    for mode in (0, 1):
        lib.pops_set_option(b"knn_fused_prepass", mode)
        res[mode] = timed(lambda: _C.knn_points_idx(p1, p2, L1, L2, 2, K, -1))
    line = (f"{name:14s} K={K:2d}: multi/fused step {res[0][2] * 1e3:7.1f} us (search {res[0][1][0] * 1e3:7.1f}, order {res[0][1][1] * 1e3:6.1f})   "
            f"cluster step {res[1][2] * 1e3:7.1f} us (search {res[1][1][0] * 1e3:7.1f}, order {res[1][1][1] * 1e3:6.1f})  equal={same(res[0][0], res[1][0])}")
    if not selfk:
        pr = {}
        for mode in (0, 1):
            lib.pops_set_option(b"knn_fused_prepass", mode)
            pr[mode] = timed(lambda: _C.knn_points_idx_pair(p1, p2, L1, L2, 2, K))
        line += (f"   pair: {pr[0][2] * 1e3:7.1f} -> {pr[1][2] * 1e3:7.1f} us (order {pr[0][1][1] * 1e3:6.1f} -> {pr[1][1][1] * 1e3:6.1f})"
                 f" equal={same(pr[0][0], pr[1][0])}")
    print(line, flush=True)
lib.pops_set_option(b"knn_fused_prepass", 1)
