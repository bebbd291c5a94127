"""Development probe: T-shape pruned search against the seed-block count and the candidate buffer capacity."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch3d_pointops_b200 import _C, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(0)
p = torch.rand(32, 16384, 3, generator=g).to(dev)
L = torch.full((32,), 16384, dtype=torch.int64, device=dev)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
def kernel_us(n=20):
    f = lambda: _C.knn_points_idx(p, p, L, L, 2, K, -1)
    for _ in range(3): f()
    torch.cuda.synchronize()
    lib.pops_profile_reset(); lib.pops_profile_enable(1)
    for _ in range(n):
        flush.zero_(); f()
    torch.cuda.synchronize()
    nl, ms = ctypes.c_int64(0), ctypes.c_double(0)
    lib.pops_profile_read(b"knn_scan", ctypes.byref(nl), ctypes.byref(ms))
    lib.pops_profile_enable(0)
    return ms.value / max(nl.value, 1) * 1e3
ref = None
for nseed in (3,):
    for cap in (24, 48, 24, 48):
        lib.pops_set_option(b"knn_nseed", nseed); lib.pops_set_option(b"knn_bufcap", cap)
        out = _C.knn_points_idx(p, p, L, L, 2, K, -1)
        if ref is None: ref = out
        same = all(torch.equal(a, b) for a, b in zip(ref, out))
        print(f"K={K} nseed={nseed} bufcap={cap}: kernel {kernel_us():7.1f} us  same={same}", flush=True)
