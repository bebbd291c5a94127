"""Development probe: a few two-sided K = 1 searches on the chamfer shape (configs[1]) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pytorch3d_pointops_b200 import _C
ch = {k: v.cuda() for k, v in bench.make_chamfer_inputs(0).items()}
for _ in range(3):
    _C.knn_points_idx_pair(ch["x"], ch["y"], ch["xl"], ch["yl"], 2, 1)
torch.cuda.synchronize()
print("ok")
