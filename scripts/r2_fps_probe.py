"""Development probe (round 2): FPS cluster exchange -- one-sided mbarrier push (fps_push=1) vs CTA winner +
cluster barrier (fps_push=0): same indices, time per iteration."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_pointops_b200 import _lib  # noqa: E402
from pytorch3d_pointops_b200.functions import sample_farthest_points  # noqa: E402

lib = _lib.load()
g = torch.Generator().manual_seed(0)
for (N, P, K) in [(8, 65536, 1024), (64, 65536, 1024), (64, 8192, 512), (4, 16384, 256), (16, 3000, 128), (2, 200000, 64)]:
    pts = torch.rand(N, P, 3, generator=g).cuda()
    L = torch.randint(P // 2, P + 1, (N,), generator=g).cuda()
    res = {}
    for push in (0, 1):
        lib.pops_set_option(b"fps_push", push)
        for _ in range(2):
            out = sample_farthest_points(pts, L, K=K)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = sample_farthest_points(pts, L, K=K); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[push] = (out, sorted(ts)[2])
    eq = torch.equal(res[0][0][1], res[1][0][1])
    print(f"N={N} P={P} K={K}: barrier {res[0][1]:.3f} ms ({res[0][1] * 1e3 / K:.2f} us/iter)   push {res[1][1]:.3f} ms "
          f"({res[1][1] * 1e3 / K:.2f} us/iter)  equal={eq}", flush=True)
lib.pops_set_option(b"fps_push", 1)
