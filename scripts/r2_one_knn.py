"""Development probe: a few T-shape KNN calls (K from argv, option knn_v2 from POPS_KNN_V2) for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_pointops_b200 import _C  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator().manual_seed(0)
p = torch.rand(32, 16384, 3, generator=g).cuda()
L = torch.full((32,), 16384, dtype=torch.int64, device="cuda")
for _ in range(3):
    _C.knn_points_idx(p, p, L, L, 2, K, -1)
torch.cuda.synchronize()
print("ok")
