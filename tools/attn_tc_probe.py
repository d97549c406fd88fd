"""A few launches of the wide-window attention forward (tcgen05) for ncu, and CUDA-graph timings of the tcgen05, mma.sync
and CUDA-core kernels at the same shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import favit_b200  # noqa: F401
from favit_b200 import _lib as L, raw
from kernel_bench import PEAK_GB, timeit_graph

bf = torch.bfloat16
for (B, N, H, W) in ((15, 4097, 6, 63), (332, 197, 12, 63), (332, 197, 12, 31), (15, 4097, 6, 15), (15, 4097, 6, 7)):
    D = H * 64
    qkv = [torch.randn(B * N, 3 * D, device="cuda").to(bf) for _ in range(3)]
    do = [torch.randn(B * N, D, device="cuda").to(bf) for _ in range(3)]
    t = timeit_graph(lambda i: raw.attn_fwd(qkv[i], B, N, H, 64, W), 8, 3)
    k = L.last_kernel().split(" ")[0]
    o, lse = raw.attn_fwd(qkv[0], B, N, H, 64, W)
    tb = timeit_graph(lambda i: raw.attn_bwd(qkv[0], o, lse, do[i], B, N, H, 64, W)[0], 4, 3)
    kb = L.last_kernel().split(" ")[0]
    by = 4.0 * B * N * D * 2
    print(f"N={N} H={H} W={W}: fwd {t:8.1f} us {by / t / 1e3 / PEAK_GB:5.3f} of HBM [{k}]   bwd {tb:8.1f} us "
          f"{2 * by / tb / 1e3 / PEAK_GB:5.3f} of HBM [{kb}]", flush=True)
