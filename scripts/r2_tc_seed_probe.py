"""Development probe: tensor-core KNN (C5 shape) scan / re-rank time against the number of seed tiles."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch3d_pointops_b200 import _C, _lib
N, P, D, K = int(os.environ.get("TC_N", 16)), 32768, 128, 16
g = torch.Generator().manual_seed(4)
x = torch.randn(N, P, D, generator=g).cuda()
L = torch.full((N,), P, device="cuda")
lib = _lib.load()
ref = None
for seed in [int(a) for a in sys.argv[1:]] or [4, 8, 16, 32]:
    lib.pops_set_option(b"tc_seed", seed)
    for _ in range(2):
        out = _C.knn_points_idx(x, x, L, L, 2, K, -1)
    torch.cuda.synchronize()
    if ref is None:
        ref = out
    same = all(torch.equal(a, b) for a, b in zip(ref, out))
    lib.pops_profile_reset(); lib.pops_profile_enable(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); _C.knn_points_idx(x, x, L, L, 2, K, -1); b.record(); torch.cuda.synchronize()
    parts = []
    for name in (b"knn_tc_scan", b"knn_tc_rerank", b"knn_exact_rows"):
        nl, ms = ctypes.c_int64(0), ctypes.c_double(0)
        lib.pops_profile_read(name, ctypes.byref(nl), ctypes.byref(ms))
        parts.append(f"{name.decode()} {ms.value:.3f}")
    lib.pops_profile_enable(0)
    print(f"tc_seed={seed:2d}: total {a.elapsed_time(b):.2f} ms  " + "  ".join(parts) + f"  same as first: {same}", flush=True)
