"""Runs the twelve GEMMs of ONE transformer block at a workload's shapes (default C4: ViT-B/16, 256 images ->
M = 50 432 tokens) exactly as favit::block_fwd / block_bwd issue them: forward qkv / proj / fc1(GELU dual store) /
fc2(+fp32 residual), dgrad fc2(GELU' + column sums) / fc1 / proj / qkv, wgrad x4 (fp32 split-K reduce-add).
Meant to be profiled:  ncu --set full -k regex:gemm_bf16 python tools/gemm_block_probe.py
Prints which kernel variant each call dispatched to (favit_last_kernel) and, without ncu, CUDA-event times."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import favit_b200  # noqa: F401
from favit_b200 import _lib as L, raw


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=256 * 197)
    ap.add_argument("--D", type=int, default=768)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--once", action="store_true", help="launch every GEMM exactly once (for ncu: one row per GEMM)")
    a = ap.parse_args()
    M, D, Hd = a.M, a.D, 4 * a.D
    bf = torch.bfloat16
    r = lambda *s: torch.randn(*s, device="cuda").to(bf)
    x, o, h = r(M, D), r(M, D), r(M, Hd)
    wqkv, wp, w1, w2 = r(3 * D, D) * 0.05, r(D, D) * 0.05, r(Hd, D) * 0.05, r(D, Hd) * 0.05
    bq, bp, b1, b2 = (torch.randn(n, device="cuda") for n in (3 * D, D, Hd, D))
    res = torch.randn(M, D, device="cuda")
    g, dq, dh, pre = r(M, D), r(M, 3 * D), r(M, Hd), r(M, Hd)
    calls = [
        ("fwd qkv", 2.0 * M * D * 3 * D, lambda: raw.linear_fwd(x, wqkv, bq, None, bf)),
        ("fwd proj", 2.0 * M * D * D, lambda: raw.linear_fwd(o, wp, bp, None, bf)),
        ("fwd fc1 gelu", 2.0 * M * D * Hd, lambda: raw.linear_fwd(x, w1, b1, None, bf, gelu=True, save_preact=True)),
        ("fwd fc2 residual", 2.0 * M * D * Hd, lambda: raw.linear_fwd(h, w2, b2, res, torch.float32)),
        ("dgrad fc2 gelu' colsum", 2.0 * M * D * Hd, lambda: raw.linear_dgrad(g, w2, pre, bf, colsum=True)),
        ("dgrad fc1", 2.0 * M * D * Hd, lambda: raw.linear_dgrad(dh, w1, None, bf)),
        ("dgrad proj", 2.0 * M * D * D, lambda: raw.linear_dgrad(g, wp, None, bf)),
        ("dgrad qkv", 2.0 * M * D * 3 * D, lambda: raw.linear_dgrad(dq, wqkv, None, bf)),
        ("wgrad fc2", 2.0 * M * D * Hd, lambda: raw.linear_wgrad(g, h, want_bias=False)),
        ("wgrad fc1", 2.0 * M * D * Hd, lambda: raw.linear_wgrad(dh, x, want_bias=False)),
        ("wgrad proj", 2.0 * M * D * D, lambda: raw.linear_wgrad(g, o, want_bias=False)),
        ("wgrad qkv", 2.0 * M * D * 3 * D, lambda: raw.linear_wgrad(dq, x, want_bias=False)),
    ]
    for name, flops, fn in calls:
        fn()
        torch.cuda.synchronize()
        kern = L.last_kernel()
        if a.once:
            print(f"{name:26s} {kern}", flush=True)
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / a.reps * 1e3
        print(f"{name:26s} {us:8.1f} us {flops / us / 1e6:8.1f} TFLOP/s  {kern}", flush=True)


if __name__ == "__main__":
    main()
