"""Synthetic inputs of the shapes BASELINE.json names (there is no network for datasets).

Label maps stand in for `skimage.segmentation.slic` output (models/sppp.py:61-68), which is an *input* of the hot
path: jittered-grid Voronoi cells, K seeds on a sqrt(K) x sqrt(K) grid, each moved by up to +-jitter of a cell.
"""
from __future__ import annotations

import math

import torch


def _dominant_label_count(lm: torch.Tensor, patch_size: int, K: int) -> torch.Tensor:
    """Number of distinct dominant labels per image (labels must lie in [0, K))."""
    B, S, _ = lm.shape
    g = S // patch_size
    t = lm[:, :g * patch_size, :g * patch_size].reshape(B, g, patch_size, g, patch_size)
    t = t.permute(0, 1, 3, 2, 4).reshape(B, g * g, patch_size * patch_size)
    hist = torch.zeros(B, g * g, K, dtype=torch.int32, device=lm.device)
    hist.scatter_add_(2, t, torch.ones_like(t, dtype=torch.int32))
    dom = hist.argmax(dim=2)
    present = torch.zeros(B, K, dtype=torch.bool, device=lm.device)
    present.scatter_(1, dom, True)
    return present.sum(dim=1)


def voronoi_label_maps(B: int, S: int, K: int, seed: int = 0, device="cuda", jitter: float = 0.35,
                       exact_k: bool = False, patch_size: int = 16) -> torch.Tensor:
    """int64 [B,S,S] label maps with labels 0..K-1.  With exact_k every image is redrawn until exactly K distinct labels
    dominate at least one patch (the reference needs R == num_superpixels, models/sppp_mhla.py:300,310)."""
    k = int(round(math.sqrt(K)))
    if k * k != K:
        raise ValueError("K must be a perfect square")
    gen = torch.Generator(device="cpu").manual_seed(seed)
    cell = S / k
    c = (torch.arange(k, dtype=torch.float32) + 0.5) * cell
    cy, cx = torch.meshgrid(c, c, indexing="ij")
    cy, cx = cy.reshape(-1).to(device), cx.reshape(-1).to(device)
    pix = (torch.arange(S, dtype=torch.float32, device=device) + 0.5)

    def draw(n):
        jy = ((torch.rand(n, K, generator=gen) * 2 - 1) * jitter * cell).to(device)
        jx = ((torch.rand(n, K, generator=gen) * 2 - 1) * jitter * cell).to(device)
        sy, sx = cy[None] + jy, cx[None] + jx                                  # [n,K]
        out = torch.empty(n, S, S, dtype=torch.int64, device=device)
        for i in range(n):                                                     # [K,S,S] distance volume per image
            d = (pix[None, :, None] - sy[i][:, None, None]) ** 2 + (pix[None, None, :] - sx[i][:, None, None]) ** 2
            out[i] = d.argmin(dim=0)
        return out

    lm = draw(B)
    if exact_k:
        for _ in range(50):
            bad = (_dominant_label_count(lm, patch_size, K) != K).nonzero().flatten()
            if bad.numel() == 0:
                break
            lm[bad] = draw(int(bad.numel()))
        else:
            raise RuntimeError("could not draw label maps with exactly K dominant labels")
    return lm


def images(B: int, S: int, seed: int = 0, device="cuda", channels: int = 3) -> torch.Tensor:
    """x ~ N(0,1) fp32 [B,C,S,S], like the reference's own benchmark input (utils/metrics.py:336)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    return torch.randn(B, channels, S, S, generator=gen, device=device)


def class_labels(B: int, num_classes: int, seed: int = 0, device="cuda") -> torch.Tensor:
    gen = torch.Generator(device=device).manual_seed(seed + 1)
    return torch.randint(0, num_classes, (B,), generator=gen, device=device)
