"""Which tcgen05 GEMM variant wins at the SPPP ViT-S shapes (M = 4352 tokens)?  Plain epilogue, bf16 or fp32 C;
bn 64/128/256 = single-CTA kernel, 512 = CTA-pair kernel, 0 = what the cost model picks.  CUDA-graph timing.
usage: python tools/gemm_small_sweep.py [M]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import favit_b200  # noqa: F401
from favit_b200 import _lib as L
from kernel_bench import timeit_graph

M0 = int(sys.argv[1]) if len(sys.argv) > 1 else 4352
D, Hd = 384, 1536
bf = torch.bfloat16
dev = "cuda"
# name, a_mn, b_mn, M, N, K, c_dtype
cases = [
    ("qkv fwd", 0, 0, M0, 3 * D, D, L.BF16), ("proj fwd", 0, 0, M0, D, D, L.BF16), ("fc1 fwd", 0, 0, M0, Hd, D, L.BF16),
    ("fc2 fwd", 0, 0, M0, D, Hd, L.F32),
    ("fc2 dgrad", 0, 1, M0, Hd, D, L.BF16), ("fc1 dgrad", 0, 1, M0, D, Hd, L.BF16), ("qkv dgrad", 0, 1, M0, D, 3 * D, L.BF16),
    ("proj dgrad", 0, 1, M0, D, D, L.BF16),
    ("fc2 wgrad", 1, 1, D, Hd, M0, L.F32), ("fc1 wgrad", 1, 1, Hd, D, M0, L.F32), ("qkv wgrad", 1, 1, 3 * D, D, M0, L.F32),
    ("proj wgrad", 1, 1, D, D, M0, L.F32),
]
st = torch.cuda.current_stream().cuda_stream
for name, a_mn, b_mn, M, N, K, cdt in cases:
    nbuf = 3
    A = [torch.randn((K, M) if a_mn else (M, K), device=dev).to(bf) for _ in range(nbuf)]
    B = [torch.randn((K, N) if b_mn else (N, K), device=dev).to(bf) for _ in range(nbuf)]
    C = torch.zeros(M, N, device=dev, dtype=torch.float32 if cdt == L.F32 else bf)
    res = []
    for bn in (0, 64, 128, 256, 512):
        for sp in ((1,) if cdt != L.F32 else (1, 2, 4, 8)):
            def run(i, bn=bn, sp=sp):
                side = torch.cuda.current_stream().cuda_stream
                rc = L.lib().favit_gemm_bf16_raw(A[i].data_ptr(), a_mn, A[i].stride(0), B[i].data_ptr(), b_mn, B[i].stride(0),
                                                 C.data_ptr(), N, cdt, M, N, K, bn, sp if bn else 0, side)
                return rc
            if run(0) != 0:
                continue
            torch.cuda.synchronize()
            try:
                us = timeit_graph(run, 10, nbuf)
            except Exception as e:  # noqa: BLE001
                continue
            res.append((us, bn, sp))
    res.sort()
    auto = [r for r in res if r[1] == 0]
    print(f"{name:11s} M{M} N{N} K{K}: auto {auto[0][0]:.1f} us | best " + ", ".join(f"bn{b}/s{s} {u:.1f}" for u, b, s in res[:4]), flush=True)
