"""Probe: where does ln_bwd spend its time?  Variants with parts switched off (C ABI called directly)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import favit_b200
from favit_b200 import _lib as L

M, D = 50432, 768
dev = "cuda"
nb = 3
x = [torch.randn(M, D, device=dev) for _ in range(nb)]
dy = [torch.randn(M, D, device=dev).to(torch.bfloat16) for _ in range(nb)]
dres = [torch.randn(M, D, device=dev) for _ in range(nb)]
mean = torch.zeros(M, device=dev); rstd = torch.ones(M, device=dev); g = torch.ones(D, device=dev)
dx = torch.empty(M, D, device=dev); dxb = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
acc = torch.zeros(3, D, device=dev)
st = torch.cuda.current_stream().cuda_stream
p = lambda t: t.data_ptr() if t is not None else None

def run(name, use_dres, use_bf, use_acc, bytes_):
    def f(i):
        rc = L.lib().favit_layernorm_bwd(p(dy[i]), L.BF16, p(x[i]), L.F32, p(mean), p(rstd), p(g), p(dres[i]) if use_dres else None,
                                         p(dx), p(dxb) if use_bf else None, acc[0].data_ptr() if use_acc else None,
                                         acc[1].data_ptr() if use_acc else None, acc[2].data_ptr() if use_acc else None, M, D, st)
        assert rc == 0
    for i in range(3): f(i % nb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): f(i % nb)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"{name:40s} {us:7.1f} us  {bytes_ / us / 1e3:7.1f} GB/s")

E = M * D
run("full (dres, bf16 copy, dgamma)", True, True, True, E * (2 + 4 + 4 + 4 + 2))
run("no dgamma/dbeta/dxsum", True, True, False, E * (2 + 4 + 4 + 4 + 2))
run("no dres", False, True, True, E * (2 + 4 + 4 + 2))
run("no bf16 copy", True, False, True, E * (2 + 4 + 4 + 4))
run("minimal (dy, x -> dx)", False, False, False, E * (2 + 4 + 4))
