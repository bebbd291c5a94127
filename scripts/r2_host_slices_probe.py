"""Development probe: host.HostKnn e2e time on the T shape against the slice schedule."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch3d_pointops_b200.host import HostKnn
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
p = torch.rand(32, 16384, 3, generator=g).pin_memory()
L = torch.full((32,), 16384, dtype=torch.int64).pin_memory()
def run(slices, n=15):
    hk = HostKnn(32, 16384, 16384, 3, 16, dev, slices=slices)
    for _ in range(3):
        hk(p, None, L, L); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); hk(p, None, L, L); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]
for s in (8, [1, 3, 4, 4, 4, 4, 4, 4, 4], [2, 2, 4, 4, 4, 4, 4, 4, 4], [1, 1, 2, 4, 4, 4, 4, 4, 4, 4], [2, 6, 8, 8, 8],
          [1, 3, 4, 8, 8, 8], [1, 7, 8, 8, 8], [2, 4, 4, 4, 6, 6, 6], 8):
    med, best = run(s)
    print(f"slices={s}: median {med:.3f} ms  best {best:.3f} ms", flush=True)
