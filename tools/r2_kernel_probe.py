"""One launch each of the kernels added in round 2, at benchmark shapes — meant for
    ncu --set full -k regex:'attn_tc|pool_pixels|pool_bwd_rows|dominant_vec|adamw|dropout_cast' python tools/r2_kernel_probe.py
Without ncu it prints CUDA-graph timings (16 launches over rotating buffers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import favit_b200  # noqa: F401
from favit_b200 import _lib as L, ops, raw, synth
from kernel_bench import PEAK_GB, timeit_graph as _timeit_graph

ONCE = "--once" in sys.argv      # for ncu: launch every kernel exactly once instead of timing it


def timeit_graph(fn, iters, nbuf):
    if ONCE:
        fn(0)
        torch.cuda.synchronize()
        return 1.0
    return _timeit_graph(fn, iters, nbuf)


def main():
    dev = "cuda"
    bf = torch.bfloat16
    out = []
    # wide-window attention forward, C5-as-ViT tokens
    for (B, N, H, W) in ((15, 4097, 6, 63), (332, 197, 12, 63), (332, 197, 12, 31)):
        D = H * 64
        qkv = [torch.randn(B * N, 3 * D, device=dev).to(bf) for _ in range(3)]
        do = [torch.randn(B * N, D, device=dev).to(bf) for _ in range(3)]
        t = timeit_graph(lambda i: raw.attn_fwd(qkv[i], B, N, H, 64, W), 8, 3)
        out.append((f"attn fwd N={N} H={H} W={W} [{L.last_kernel().split(' ')[0]}]", t, 4.0 * B * N * D * 2))
        o, lse = raw.attn_fwd(qkv[0], B, N, H, 64, W) if not ONCE else (torch.empty(B * N, D, device=dev, dtype=bf), torch.zeros(B, H, N, device=dev))
        t = timeit_graph(lambda i: raw.attn_bwd(qkv[0], o, lse, do[i], B, N, H, 64, W)[0], 4, 3)
        out.append((f"attn bwd N={N} H={H} W={W} [{L.last_kernel().split(' ')[0]}]", t, 8.0 * B * N * D * 2))
    # SPPP at C2
    B, S, ps, K, D = 256, 224, 16, 16, 384
    lms = [synth.voronoi_label_maps(B, S, K, seed=3 + i, device=dev, exact_k=True, patch_size=ps) for i in range(3)]
    asg = [ops.sppp_assign(lm, ps, S, K) for lm in lms]
    img = [torch.randn(B, 3, S, S, device=dev) for _ in range(3)]
    t = timeit_graph(lambda i: ops.sppp_pool_pixels(img[i], asg[i][6], asg[i][5], asg[i][2], ps, K, bf), 8, 3)
    out.append((f"sppp_pool_pixels C2 [{L.last_kernel()}]", t, B * 3 * S * S * 4.0 + B * K * 768 * 2))
    t = timeit_graph(lambda i: ops.sppp_assign_centroids(lms[i], ps, S, K, K), 8, 3)
    out.append((f"sppp_assign_centroids C2 [{L.last_kernel()}]", t, B * S * S * 8.0))
    t = timeit_graph(lambda i: ops.sppp_assign(lms[i], ps, S, K), 8, 3)
    out.append(("sppp_assign C2 (label pass without centroids)", t, B * S * S * 8.0))
    t = timeit_graph(lambda i: ops.sppp_centroids(lms[i], K), 8, 3)
    out.append(("sppp_centroids C2 (separate pass)", t, B * S * S * 8.0))
    g = [torch.randn(B, K, D, device=dev) for _ in range(3)]
    t = timeit_graph(lambda i: ops.sppp_pool_bwd(g[i], asg[i][1], asg[i][3], bf), 8, 3)
    out.append((f"sppp_pool_bwd C2 [{L.last_kernel()}]", t, B * K * D * 4.0 + B * 196 * 4 + B * 196 * D * 2.0))
    # fused dropout GEMMs at C4 and the masked operand copy
    M, Dm, Hd = 256 * 197, 768, 3072
    x, w1 = torch.randn(M, Dm, device=dev).to(bf), (torch.randn(Hd, Dm, device=dev) * 0.05).to(bf)
    b1 = torch.randn(Hd, device=dev)
    seed = torch.tensor([7], dtype=torch.int64, device=dev)
    for drop in (None, (0.1, seed, raw.drop_offset(0, 1))):
        t = timeit_graph(lambda i: raw.linear_fwd(x, w1, b1, None, bf, gelu=True, save_preact=True, drop=drop), 6, 1)
        out.append((f"fc1 fwd GELU {'+ dropout' if drop else '(no dropout)'} C4", t, None, 2.0 * M * Dm * Hd))
    gg = torch.randn(M, Dm, device=dev)
    t = timeit_graph(lambda i: raw.dropout_cast(gg, bf, (0.1, seed, raw.drop_offset(0, 2)), colsum=torch.zeros(Dm, device=dev)), 6, 1)
    out.append(("dropout_cast C4", t, M * Dm * 6.0))
    # AdamW over ViT-B's parameters
    from favit_b200.models import VisionTransformerMHLA
    from favit_b200.optim import FusedAdamW
    m = VisionTransformerMHLA(img_size=224, patch_size=16, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                              window_size=7, use_mhla=True).to(dev)
    for p in m.parameters():
        p.grad = torch.randn_like(p)
    n = sum(p.numel() for p in m.parameters())
    for name, opt in (("favit FusedAdamW", FusedAdamW(m.parameters(), lr=1e-4, weight_decay=0.05)),
                      ("torch AdamW(fused)", torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=0.05, fused=True, capturable=True))):
        opt.step()
        t = timeit_graph(lambda i: opt.step(), 4, 1)
        out.append((f"{name}, 86.6 M parameters", t, n * 28.0))
    for rec in out:
        name, us = rec[0], rec[1]
        if rec[2] is not None:
            print(f"{name:90s} {us:9.1f} us  {rec[2] / us / 1e3:8.1f} GB/s = {rec[2] / us / 1e3 / PEAK_GB:5.3f} of HBM", flush=True)
        else:
            print(f"{name:90s} {us:9.1f} us  {rec[3] / us / 1e6:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
