"""Development aid: per-kernel timing of the tensor-core KNN path on the C5 shape."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch3d_pointops_b200 import _C, _lib
N = int(os.environ.get("TC_N", 16)); P = int(os.environ.get("TC_P", 32768)); D = int(os.environ.get("TC_D", 128)); K = 16
g = torch.Generator().manual_seed(4)
x = torch.randn(N, P, D, generator=g).cuda()
L = torch.full((N,), P, device="cuda")
lib = _lib.load()
for _ in range(2):
    _C.knn_points_idx(x, x, L, L, 2, K, -1)
torch.cuda.synchronize()
lib.pops_profile_reset(); lib.pops_profile_enable(1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); _C.knn_points_idx(x, x, L, L, 2, K, -1); b.record(); torch.cuda.synchronize()
print(f"total {a.elapsed_time(b):.2f} ms")
for name in (b"knn_tc_scan", b"knn_tc_rerank", b"knn_exact_rows", b"knn_generic"):
    nl, ms = ctypes.c_int64(0), ctypes.c_double(0)
    lib.pops_profile_read(name, ctypes.byref(nl), ctypes.byref(ms))
    print(name.decode(), nl.value, f"{ms.value:.3f} ms")
lib.pops_profile_enable(0)
