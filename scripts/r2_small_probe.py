"""Development probe: where the time of a small knn_points call goes (launch-bound regime of the
reference's timing harness: B = 1..32 clouds of 100..2000 points, K = 16)."""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_pointops_b200 import _C, _lib  # noqa: E402
from pytorch3d_pointops_b200.functions import knn_points  # noqa: E402

lib = _lib.load()


def kernel_ms(name):
    n_, ms_ = ctypes.c_int64(0), ctypes.c_double(0.0)
    lib.pops_profile_read(name, ctypes.byref(n_), ctypes.byref(ms_))
    return ms_.value / max(1, n_.value)


for opt in (-1, 1):
    lib.pops_set_option(b"knn_order", opt)
    for B, P in ((1, 100), (1, 500), (32, 500), (1, 1000), (1, 2000), (8, 2000)):
        x = torch.randn(B, P, 3, device="cuda")
        for _ in range(5):
            knn_points(x, x, K=16)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        t0 = time.perf_counter()
        for _ in range(20):
            knn_points(x, x, K=16)
            torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 20 * 1e3
        launches = (_lib.launch_count() - l0) / 20
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            knn_points(x, x, K=16)
        b.record()
        torch.cuda.synchronize()
        stream_ms = a.elapsed_time(b) / 20
        lib.pops_profile_reset()
        lib.pops_profile_enable(1)
        for _ in range(5):
            knn_points(x, x, K=16)
        torch.cuda.synchronize()
        lib.pops_profile_enable(0)
        print(f"knn_order={opt:2d} B={B:2d} P={P:5d}: wall+sync {wall:.3f} ms, back-to-back {stream_ms:.3f} ms, "
              f"launches {launches:.0f}, knn_scan kernel {kernel_ms(b'knn_scan') * 1e3:.1f} us", flush=True)
        lib.pops_profile_reset()
lib.pops_set_option(b"knn_order", -1)
