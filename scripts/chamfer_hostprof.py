import sys, os, cProfile, pstats
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
for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr).sort_stats("cumulative")
st.print_stats(28)
