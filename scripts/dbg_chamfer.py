import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from conftest import Golden
from test_oracle import run_chamfer_variant
import ast
from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance
g=Golden("chamfer_cases")
vs=[ast.literal_eval(str(s)) for s in g.a("variants")]
for vi,v in enumerate(vs):
    flat,grads=run_chamfer_variant(chamfer_distance,g,v,"cuda:0")
    for i,t in enumerate(flat):
        w=g.t(f"v{vi}.out{i}"); d=(t.detach().cpu()-w).abs()
        rel=(d/(w.abs()+1e-30)).max().item()
        if not torch.allclose(t.detach().cpu(),w,rtol=1e-5,atol=1e-7): print("OUT",vi,i,v,"maxabs",d.max().item(),"maxrel",rel, "shape",tuple(w.shape))
    for n,gr in grads.items():
        w=g.t(f"v{vi}.g_{n}")
        if w.numel()==0: continue
        d=(gr.cpu()-w).abs()
        if not torch.allclose(gr.cpu(),w,rtol=1e-5,atol=1e-5*float(w.abs().max())): print("GRAD",vi,n,"maxabs",d.max().item(),"scale",w.abs().max().item())
