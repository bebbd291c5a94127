import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch3d_pointops_b200.functions import sample_farthest_points
g = torch.Generator().manual_seed(0)
for (N, P, K) in [(64, 4096, 512), (64, 8192, 512), (8, 65536, 1024), (64, 65536, 1024), (18, 65536, 1024)]:
    pts = torch.rand(N, P, 3, generator=g).cuda()
    for _ in range(2): sample_farthest_points(pts, K=K)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); sample_farthest_points(pts, K=K); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"C={os.environ.get('POPS_FPS_C','auto')} N={N} P={P} K={K}: {ms:.3f} ms  {ms*1e3/K:.2f} us/iter")
