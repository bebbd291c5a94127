import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance
dev = torch.device("cuda:0")
ch = {k: v.to(dev) for k, v in bench.make_chamfer_inputs(0).items()}
for k in ("x", "y", "xn", "yn", "xc", "yc"):
    ch[k].requires_grad_(True)
def step():
    for k in ("x", "y", "xn", "yn", "xc", "yc"):
        ch[k].grad = None
    loss, lf = chamfer_distance(ch["x"], ch["y"], x_lengths=ch["xl"], y_lengths=ch["yl"],
                                x_features={"normals": ch["xn"], "colors": ch["xc"]},
                                y_features={"normals": ch["yn"], "colors": ch["yc"]},
                                feature_names=["normals", "colors"])
    (loss + lf["normals"] + lf["colors"]).backward()
for _ in range(4):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(20):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per step {(t1 - t0) / 20 * 1e3:.3f} ms, incl. device drain {(t2 - t0) / 20 * 1e3:.3f} ms")
