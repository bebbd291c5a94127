"""Development probe: a few FPS calls on 8 clouds of 65536 points (K=1024), for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_pointops_b200.functions import sample_farthest_points  # noqa: E402

g = torch.Generator().manual_seed(0)
pts = torch.rand(8, 65536, 3, generator=g).cuda()
for _ in range(3):
    sample_farthest_points(pts, K=1024)
torch.cuda.synchronize()
print("ok")
