import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, favit_b200
from favit_b200 import raw
from favit_b200.mhla import fold_latent
D, H = 768, 12
dev = "cuda"
qw, qb = torch.randn(3*D, D, device=dev), torch.randn(3*D, device=dev)
pw, pb = torch.randn(D, D, device=dev), torch.randn(D, device=dev)
lw, lb = torch.randn(64, 64, device=dev), torch.randn(64, device=dev)
def t(fn, n=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("fold_fwd us", t(lambda: raw.fold_fwd(qw, qb, pw, pb, lw, lb, H, torch.bfloat16)))
dwq, dbq, dwp, dbp = torch.randn(3*D, D, device=dev), torch.randn(3*D, device=dev), torch.randn(D, D, device=dev), torch.randn(D, device=dev)
print("fold_bwd us", t(lambda: raw.fold_bwd(qw, qb, pw, lw, lb, dwq, dbq, dwp, dbp, H)))
print("torch fold fwd us", t(lambda: fold_latent(qw, qb, pw, pb, lw, lb, H)))
