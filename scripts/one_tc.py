import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch3d_pointops_b200 import _C
N, P, D, K = int(os.environ.get("TC_N", 4)), 32768, 128, 16
g = torch.Generator().manual_seed(4)
x = torch.randn(N, P, D, generator=g).cuda()
L = torch.full((N,), P, device="cuda")
for _ in range(2):
    _C.knn_points_idx(x, x, L, L, 2, K, -1)
torch.cuda.synchronize()
