"""One warmed-up invocation of the secondary ops (profiling target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pytorch3d_pointops_b200.functions import ball_query, sample_farthest_points, knn_points, knn_gather
from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance
dev = torch.device("cuda:0")
which = sys.argv[1]
g = torch.Generator().manual_seed(3)
if which == "bq":
    p = torch.rand(32, 16384, 3, generator=g).to(dev)
    for _ in range(3):
        ball_query(p, p, K=32, radius=0.1)
elif which == "fps":
    p = torch.rand(8, 65536, 3, generator=g).to(dev)
    for _ in range(3):
        sample_farthest_points(p, K=1024)
elif which == "chamfer":
    ch = {k: v.to(dev) for k, v in bench.make_chamfer_inputs(0).items()}
    for k in ("x", "y", "xn", "yn", "xc", "yc"):
        ch[k].requires_grad_(True)
    for _ in range(3):
        loss, lf = chamfer_distance(ch["x"], ch["y"], x_lengths=ch["xl"], y_lengths=ch["yl"],
                                    x_features={"normals": ch["xn"], "colors": ch["xc"]},
                                    y_features={"normals": ch["yn"], "colors": ch["yc"]},
                                    feature_names=["normals", "colors"])
        (loss + lf["normals"] + lf["colors"]).backward()
elif which == "knnbwd":
    p = torch.rand(32, 16384, 3, generator=g).to(dev).requires_grad_(True)
    for _ in range(3):
        r = knn_points(p, p, K=16, return_nn=True)
        (r.dists.sum() + r.knn.sum()).backward()
torch.cuda.synchronize()
