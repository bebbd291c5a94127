"""Development aid: block counters of the pruned search on the chamfer shape (x != y, ragged, K=1)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pytorch3d_pointops_b200 import _C, _lib
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ch = {k: v.cuda() for k, v in bench.make_chamfer_inputs(0).items()}
lib = _lib.load()
lib.pops_set_option(b"knn_stats", 1)
out = (ctypes.c_ulonglong * 8)()
lib.pops_knn_debug_stats(out)
_C.knn_points_idx(ch["x"], ch["y"], ch["xl"], ch["yl"], 2, K, -1)
lib.pops_knn_debug_stats(out)
f, s, fl, cg, ne, w = [int(x) for x in out[:6]]
print(f"K={K} warps {w}: blocks fetched/warp {f/w:.1f} scanned/warp {s/w:.1f}; flush rounds/warp {fl/w:.1f}; "
      f"non-empty slot flushes/warp {ne/w:.1f}; buffered groups/query {cg/int(ch['xl'].sum()):.1f}")
lib.pops_set_option(b"knn_stats", 0)
for _ in range(3):
    _C.knn_points_idx(ch["x"], ch["y"], ch["xl"], ch["yl"], 2, K, -1)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    _C.knn_points_idx(ch["x"], ch["y"], ch["xl"], ch["yl"], 2, K, -1)
b.record(); torch.cuda.synchronize()
print(f"x->y K={K}: {a.elapsed_time(b) / 10 * 1e3:.1f} us per call")
