"""Development probe: pruned search with sub-box tests -- kernel time and walk counters per K on the T
shape (uniform and ragged) and the chamfer pair, per-query (knn_subq=1) vs warp-box test for K > 4,
and bit-equality of the two against each other and against the brute-force order (knn_prune=0)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pytorch3d_pointops_b200 import _C, _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return sorted(ms)[len(ms) // 2] * 1e3


def stats(fn):
    out = (ctypes.c_ulonglong * 8)()
    lib.pops_set_option(b"knn_stats", 1)
    lib.pops_knn_debug_stats(out)
    fn()
    lib.pops_knn_debug_stats(out)
    lib.pops_set_option(b"knn_stats", 0)
    f, s, fl, cg, ne, w, sb = [int(x) for x in out[:7]]
    w = max(w, 1)
    return f"fetched/warp {f/w:.1f} scanned/warp {s/w:.1f} sub-boxes/warp {sb/w:.1f} flush rounds/warp {fl/w:.2f} groups/warp {cg/w:.0f}"


g = torch.Generator().manual_seed(0)
p = torch.rand(32, 16384, 3, generator=g).to(dev)
L = torch.full((32,), 16384, dtype=torch.int64, device=dev)
Lr = torch.randint(8192, 16385, (32,), generator=g).to(dev)
ch = {k: v.to(dev) for k, v in bench.make_chamfer_inputs(0).items()}

for K in (16, 8, 4, 1, 32):
    res = {}
    for subq in (0, 1):
        lib.pops_set_option(b"knn_subq", subq)
        f = lambda: _C.knn_points_idx(p, p, L, L, 2, K, -1)  # noqa: E731
        res[subq] = f()
        print(f"K={K:2d} subq={subq}: {timeit(f):7.1f} us  [{stats(f)}]", flush=True)
        if K <= 4:
            break
    if 1 in res:
        print("   subq 0 == 1:", all(torch.equal(a, b) for a, b in zip(res[0], res[1])))
    if K in (16, 1):
        lib.pops_set_option(b"knn_prune", 0)
        ref = _C.knn_points_idx(p, p, L, L, 2, K, -1)
        lib.pops_set_option(b"knn_prune", 1)
        print("   pruned == unpruned:", all(torch.equal(a, b) for a, b in zip(res[0], ref)))
for subq in (0, 1):
    lib.pops_set_option(b"knn_subq", subq)
    f = lambda: _C.knn_points_idx(p, p, Lr, Lr, 2, 16, -1)  # noqa: E731
    print(f"ragged K=16 subq={subq}: {timeit(f):7.1f} us  [{stats(f)}]", flush=True)
lib.pops_set_option(b"knn_subq", 0)
f = lambda: _C.knn_points_idx_pair(ch["x"], ch["y"], ch["xl"], ch["yl"], 2, 1)  # noqa: E731
print(f"chamfer pair K=1: {timeit(f):7.1f} us  [{stats(f)}]")
