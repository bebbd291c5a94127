"""Development probe: host cost of one chamfer_distance fwd+bwd step (tiny clouds: the device work is negligible),
wall clock per step and the cProfile top of the Python side."""
import cProfile, os, pstats, sys, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
N, P = 32, 128
x = torch.rand(N, P, 3, generator=g).to(dev).requires_grad_(True)
y = torch.rand(N, P, 3, generator=g).to(dev).requires_grad_(True)
xn = torch.randn(N, P, 3, generator=g).to(dev).requires_grad_(True)
yn = torch.randn(N, P, 3, generator=g).to(dev).requires_grad_(True)
xl = torch.full((N,), P, dtype=torch.int64, device=dev)
yl = torch.full((N,), P, dtype=torch.int64, device=dev)
def step():
    for t in (x, y, xn, yn):
        t.grad = None
    loss, lf = chamfer_distance(x, y, x_lengths=xl, y_lengths=yl, x_features={"normals": xn}, y_features={"normals": yn},
                                feature_names=["normals"])
    (loss + lf["normals"]).backward()
for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 300
for _ in range(n):
    step()
torch.cuda.synchronize()
print(f"host-bound step: {(time.perf_counter() - t0) / n * 1e3:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:5000])
