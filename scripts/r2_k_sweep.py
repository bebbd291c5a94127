"""Development probe: pruned-search kernel time (library hook) per K on the T shape, uniform and ragged."""
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
Lr = torch.randint(8192, 16385, (32,), generator=g).to(dev)
def kernel_us(f, n=20):
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
out = []
for K in (1, 4, 8, 16, 32):
    out.append(f"K={K}: {kernel_us(lambda: _C.knn_points_idx(p, p, L, L, 2, K, -1)):.1f}")
out.append(f"ragged K=16: {kernel_us(lambda: _C.knn_points_idx(p, p, Lr, Lr, 2, 16, -1)):.1f}")
print("kernel us  " + "  ".join(out))
