"""Development probe: where the tensor-core scan's time goes (option tc_dbg: 1 = no candidate is ever
buffered, 2 = the epilogue drains nothing (MMA / TMA pipeline alone), 4 = accumulators read, nothing
evaluated).  Results under tc_dbg != 0 are not valid searches."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch3d_pointops_b200 import _C, _lib
N, P, D, K = int(os.environ.get("TC_N", 16)), 32768, 128, 16
g = torch.Generator().manual_seed(4)
x = torch.randn(N, P, D, generator=g).cuda()
L = torch.full((N,), P, device="cuda")
lib = _lib.load()
for dbg in (0, 1, 4, 2):
    lib.pops_set_option(b"tc_dbg", dbg)
    for _ in range(2):
        _C.knn_points_idx(x, x, L, L, 2, K, -1)
    torch.cuda.synchronize()
    lib.pops_profile_reset(); lib.pops_profile_enable(1)
    _C.knn_points_idx(x, x, L, L, 2, K, -1); torch.cuda.synchronize()
    nl, ms = ctypes.c_int64(0), ctypes.c_double(0)
    lib.pops_profile_read(b"knn_tc_scan", ctypes.byref(nl), ctypes.byref(ms))
    lib.pops_profile_enable(0)
    print(f"tc_dbg={dbg}: knn_tc_scan {ms.value:.3f} ms", flush=True)
lib.pops_set_option(b"tc_dbg", 0)
