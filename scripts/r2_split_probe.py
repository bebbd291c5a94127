"""Development probe (round 2): split search (knn_split=1) vs the one-kernel search on the T shape."""
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


def run(p, L, K, split, reps=10):
    lib.pops_set_option(b"knn_split", split)
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
    lib.pops_set_option(b"knn_stats", 1)
    buf = (ctypes.c_ulonglong * 8)()
    lib.pops_knn_debug_stats(buf)
    _C.knn_points_idx(p, p, L, L, 2, K, -1)
    lib.pops_knn_debug_stats(buf)
    lib.pops_set_option(b"knn_stats", 0)
    lib.pops_set_option(b"knn_split", 1)
    return out, ms, step, int(buf[6])


g = torch.Generator().manual_seed(0)
p = torch.rand(32, 16384, 3, generator=g).to(DEV)
L = torch.full((32,), 16384, dtype=torch.int64, device=DEV)
Lr = torch.randint(8192, 16385, (32,), generator=g).to(DEV)
for name, LL in (("uniform", L), ("ragged", Lr)):
    for K in (16, 12, 8, 5):
        (ri, rd), ms0, st0, _ = run(p, LL, K, 0)
        (i, d), ms1, st1, ovf = run(p, LL, K, 1)
        print(f"{name} K={K:2d}: one-kernel {ms0 * 1e3:7.1f} us (step {st0 * 1e3:7.1f})   split {ms1 * 1e3:7.1f} us (step {st1 * 1e3:7.1f})  "
              f"equal={torch.equal(i, ri) and torch.equal(d, rd)} overflowed queries={ovf}", flush=True)
