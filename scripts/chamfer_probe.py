import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance
from pytorch3d_pointops_b200 import _C
dev = "cuda:0"
g = torch.Generator().manual_seed(1)
N, P = 32, 8192
x = torch.rand(N, P, 3, generator=g).to(dev).requires_grad_(True)
y = torch.rand(N, P, 3, generator=g).to(dev).requires_grad_(True)
xl = torch.randint(4096, P + 1, (N,), generator=g).to(dev)
yl = torch.randint(4096, P + 1, (N,), generator=g).to(dev)
xn = torch.nn.functional.normalize(torch.randn(N, P, 3, generator=g), dim=-1).to(dev).requires_grad_(True)
yn = torch.nn.functional.normalize(torch.randn(N, P, 3, generator=g), dim=-1).to(dev).requires_grad_(True)
xc = torch.rand(N, P, 3, generator=g).to(dev).requires_grad_(True)
yc = torch.rand(N, P, 3, generator=g).to(dev).requires_grad_(True)
def step():
    loss, lf = chamfer_distance(x, y, x_lengths=xl, y_lengths=yl, x_features={"normals": xn, "colors": xc},
                                y_features={"normals": yn, "colors": yc}, feature_names=["normals", "colors"])
    (loss + lf["normals"] + lf["colors"]).backward()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
import time
t0=time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
print("wall per step ms", (time.perf_counter()-t0)/20*1e3)
