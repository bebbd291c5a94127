import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch3d_pointops_b200 import _C
N, P, K = 32, 16384, int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator().manual_seed(0)
p = torch.rand(N, P, 3, generator=g).cuda()
L = torch.full((N,), P, device="cuda")
for _ in range(3):
    _C.knn_points_idx(p, p, L, L, 2, K, -1)
torch.cuda.synchronize()
