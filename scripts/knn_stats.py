"""Development aid: block / flush counters of the pruned KNN kernel (POPS_KNN_STATS=1)."""
import ctypes, os, sys
os.environ["POPS_KNN_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch3d_pointops_b200 import _C, _lib
N, P = 32, 16384
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator().manual_seed(0)
p = torch.rand(N, P, 3, generator=g).cuda()
L = torch.full((N,), P, device="cuda")
lib = _lib.load()
out = (ctypes.c_ulonglong * 8)()
lib.pops_knn_debug_stats(out)
_C.knn_points_idx(p, p, L, L, 2, K, -1)
lib.pops_knn_debug_stats(out)
f, s, fl, cg, ne, w = [int(x) for x in out[:6]]
nblk = P // 64
print(f"K={K} warps {w}: blocks fetched/warp {f/w:.1f} scanned/warp {s/w:.1f} of {nblk}; flush rounds/warp {fl/w:.1f}; "
      f"non-empty slot flushes/warp {ne/w:.1f}; buffered groups/query {cg/(N*P):.1f}")
