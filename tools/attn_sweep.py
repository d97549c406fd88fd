"""BASELINE configs[2]: MHLA attention microbench sweep — sequence lengths 17-4097 (latent tokens 16-256 + cls, patch
tokens 64-4096 + cls), heads 3/6/12, head_dim 64, window 7, bf16; B chosen so that B*N ~ 64k tokens.  Each kernel is
timed in a CUDA graph of `iters` launches over 3 rotating inputs (CUDA events).  Writes a markdown table.

    python tools/attn_sweep.py [--out gpurun_out/attn_sweep.md]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import favit_b200  # noqa: F401
from favit_b200 import raw
from kernel_bench import PEAK_GB, timeit_graph


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/attn_sweep.md")
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    bf = torch.bfloat16
    rows = []
    for N in (17, 65, 129, 197, 257, 1025, 4097):
        for H in (3, 6, 12):
            hd, W = 64, 7
            D = H * hd
            B = max(1, 65536 // N)
            M = B * N
            nbuf = 3
            qkv = [torch.randn(M, 3 * D, device="cuda").to(bf) for _ in range(nbuf)]
            do = [torch.randn(M, D, device="cuda").to(bf) for _ in range(nbuf)]
            fwd = timeit_graph(lambda i: raw.attn_fwd(qkv[i], B, N, H, hd, W), a.iters, nbuf)
            o, lse = raw.attn_fwd(qkv[0], B, N, H, hd, W)
            bwd = timeit_graph(lambda i: raw.attn_bwd(qkv[0], o, lse, do[i], B, N, H, hd, W)[0], a.iters, nbuf)
            fb, bb = 4.0 * M * D * 2, 8.0 * M * D * 2
            rows.append(dict(N=N, H=H, B=B, fwd_us=round(fwd, 1), fwd_gbps=round(fb / fwd / 1e3), bwd_us=round(bwd, 1),
                             bwd_gbps=round(bb / bwd / 1e3)))
            print(json.dumps(rows[-1]), flush=True)
            del qkv, do, o, lse
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        f.write("# MHLA window-attention sweep (bf16, head_dim 64, window 7, B*N ~ 64k tokens; CUDA-graph timing)\n\n")
        f.write(f"Algorithmic bytes: fwd 4*B*N*D*2, bwd 8*B*N*D*2; fractions are of the measured {PEAK_GB:.1f} GB/s.\n"
                "N <= 400 runs the whole-sequence TMA kernels, longer sequences the per-warp staging kernels.\n\n")
        f.write("| N | heads | B | fwd us | fwd GB/s | fwd % HBM | bwd us | bwd GB/s | bwd % HBM |\n|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write(f"| {r['N']} | {r['H']} | {r['B']} | {r['fwd_us']} | {r['fwd_gbps']} | {100 * r['fwd_gbps'] / PEAK_GB:.0f} % | "
                    f"{r['bwd_us']} | {r['bwd_gbps']} | {100 * r['bwd_gbps'] / PEAK_GB:.0f} % |\n")


if __name__ == "__main__":
    main()
