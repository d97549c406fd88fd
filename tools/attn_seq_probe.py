"""Whole-sequence window attention (mhla_window_attn_seq.cu) at the headline shapes: forward / backward time in a CUDA
graph over rotating buffers, for several L2 prefetch distances (FAVIT_SEQ_AHEAD; -100 = the resident CTA count)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import favit_b200  # noqa: F401
from favit_b200 import _lib as L, raw
from kernel_bench import PEAK_GB, timeit_graph

bf = torch.bfloat16
aheads = sys.argv[1:] or ["0", "-50", "-100", "-200"]
SHAPES = ((256, 197, 12, 7), (256, 197, 6, 7), (1024, 65, 3, 7), (63, 1025, 3, 7), (63, 1025, 12, 7), (63, 1025, 12, 15),
          (15, 4097, 3, 7), (15, 4097, 12, 7))
if os.environ.get("PROBE_SHAPES"):      # "B,N,H,W;B,N,H,W;..."
    SHAPES = tuple(tuple(int(x) for x in s.split(",")) for s in os.environ["PROBE_SHAPES"].split(";"))
for (B, N, H, W) in SHAPES:
    D = H * 64
    qkv = [torch.randn(B * N, 3 * D, device="cuda").to(bf) for _ in range(4)]
    do = [torch.randn(B * N, D, device="cuda").to(bf) for _ in range(4)]
    by = 4.0 * B * N * D * 2
    for a in aheads:
        os.environ["FAVIT_SEQ_AHEAD"] = a
        t = timeit_graph(lambda i: raw.attn_fwd(qkv[i], B, N, H, 64, W), 8, 4)
        k = L.last_kernel().split(" ")[0]
        o, lse = raw.attn_fwd(qkv[0], B, N, H, 64, W)
        tb = timeit_graph(lambda i: raw.attn_bwd(qkv[i], o, lse, do[i], B, N, H, 64, W)[0], 8, 4)
        kb = L.last_kernel().split(" ")[0]
        print(f"B={B} N={N} H={H} W={W} ahead={a:>5}: fwd {t:7.1f} us {by / t / 1e3 / PEAK_GB:5.3f} of HBM [{k}]   bwd {tb:7.1f} us "
              f"{2 * by / tb / 1e3 / PEAK_GB:5.3f} of HBM [{kb}]", flush=True)
