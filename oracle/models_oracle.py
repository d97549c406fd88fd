"""CPU oracle for the two model forwards the hot path sits in.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Functional restatements (parameters come in as a state_dict with the reference's key names) of
  * VisionTransformerMHLA.forward      -> /root/reference/models/vit_mhla.py:213-259, block :77-109,
                                          PatchEmbedding models/vit.py:36-53, MLP models/vit.py:125-139
  * SPPPViTMHLA.forward                -> /root/reference/models/sppp_mhla.py:264-325 (label maps are an input:
                                          `segmentation.segment` is SLIC, upstream of the path), centroids :226-262,
                                          DynamicPositionalEncoding models/sppp.py:271-299
The MHLA module inside uses `mhla_forward_gather` (the reference's own order of operations); SPPP uses the literal
`map_patches_oracle` + `pool_mean_oracle` per image.  bench.py times these as the CPU baseline ("port").

Parity pin: tests/golden/models.npz (outputs, loss and every parameter gradient of the real reference models).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .mhla_oracle import mhla_forward_gather
from .sppp_oracle import map_patches_oracle, pool_mean_oracle


def patch_embed(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, ps: int) -> torch.Tensor:
    """einops 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' followed by Linear (models/vit.py:36-41)."""
    B, C, H, W = x.shape
    gh, gw = H // ps, W // ps
    t = x.reshape(B, C, gh, ps, gw, ps).permute(0, 2, 4, 3, 5, 1).reshape(B, gh * gw, ps * ps * C)
    return F.linear(t, w, b)


def block(x: torch.Tensor, sd: Dict[str, torch.Tensor], pre: str, num_heads: int, window: int,
          mlp_dropout: float = 0.0, masks=None) -> torch.Tensor:
    """TransformerBlock.forward with use_mhla=True (models/vit_mhla.py:77-109).  mlp_dropout > 0 = training mode of
    models/vit.py:125-139 (dropout after the activation and after fc2) with torch's own generator: only the
    distribution is the reference's, so parity tests keep it at 0 and only the main.py-style CPU timing arm uses it.
    masks = (keep1 [B*N, hidden], keep2 [B*N, D], inv_keep): the two dropouts with EXPLICIT keep-masks (those of favit's
    counter-based generator, oracle.mlp_dropout_keep_mask), for exact parity of the fused dropout epilogues."""
    D = x.shape[-1]
    xn = F.layer_norm(x, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
    a = mhla_forward_gather(xn, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"], sd[pre + "attn.proj.weight"],
                            sd[pre + "attn.proj.bias"], sd[pre + "attn.latent_proj.weight"],
                            sd[pre + "attn.latent_proj.bias"], num_heads, window)
    x = x + a
    xn = F.layer_norm(x, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
    if pre + "mlp.fc1.weight" in sd:       # models.vit.MLP
        k1, k2 = "mlp.fc1.", "mlp.fc2."
    else:                                  # nn.Sequential of MHLATransformerBlock (models/mhla.py:197-203)
        k1, k2 = "mlp.0.", "mlp.3."
    if masks is not None:
        k1m, k2m, inv = masks
        B, N = x.shape[0], x.shape[1]
        h = F.gelu(F.linear(xn, sd[pre + k1 + "weight"], sd[pre + k1 + "bias"])) * (k1m.view(B, N, -1).to(x.dtype) * inv)
        return x + F.linear(h, sd[pre + k2 + "weight"], sd[pre + k2 + "bias"]) * (k2m.view(B, N, -1).to(x.dtype) * inv)
    h = F.dropout(F.gelu(F.linear(xn, sd[pre + k1 + "weight"], sd[pre + k1 + "bias"])), mlp_dropout, mlp_dropout > 0)
    return x + F.dropout(F.linear(h, sd[pre + k2 + "weight"], sd[pre + k2 + "bias"]), mlp_dropout, mlp_dropout > 0)


def _depth(sd) -> int:
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))


def vit_mhla_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], patch_size: int, num_heads: int,
                     window: int, mlp_dropout: float = 0.0) -> torch.Tensor:
    B = x.shape[0]
    t = patch_embed(x, sd["patch_embed.projection.1.weight"], sd["patch_embed.projection.1.bias"], patch_size)
    t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1) + sd["pos_embed"]
    for i in range(_depth(sd)):
        t = block(t, sd, f"blocks.{i}.", num_heads, window, mlp_dropout)
    D = t.shape[-1]
    t = F.layer_norm(t, (D,), sd["norm.weight"], sd["norm.bias"])
    return F.linear(t[:, 0], sd["head.weight"], sd["head.bias"])


def superpixel_centroids(seg: torch.Tensor, num_superpixels: int) -> torch.Tensor:
    """models/sppp_mhla.py:226-262: centroid (x, y) of label s in normalised coordinates, (0.5, 0.5) if absent."""
    B, H, W = seg.shape
    out = torch.zeros(B, num_superpixels, 2)
    ys = (torch.arange(H).float() / H)[:, None].expand(H, W)
    xs = (torch.arange(W).float() / W)[None, :].expand(H, W)
    for b in range(B):
        for s in range(num_superpixels):
            m = (seg[b] == s).float()
            n = m.sum()
            if n > 0:
                out[b, s, 0] = (xs * m).sum() / n
                out[b, s, 1] = (ys * m).sum() / n
            else:
                out[b, s, 0] = 0.5
                out[b, s, 1] = 0.5
    return out


def dynamic_positional_encoding(x: torch.Tensor, centroids: torch.Tensor) -> torch.Tensor:
    """models/sppp.py:271-299 (centroid branch; dropout 0)."""
    B, N, D = x.shape
    c = centroids
    if c.shape[1] < N:
        c = torch.cat([torch.ones(B, 1, 2) * 0.5, c], dim=1)
    freq = torch.exp(torch.arange(0, D // 2, dtype=torch.float) * (-math.log(10000.0) / (D // 2)))
    pe = torch.cat([torch.sin(c[:, :, 0:1] * freq), torch.cos(c[:, :, 1:2] * freq)], dim=-1)
    return x + pe


def sppp_vit_mhla_forward(x: torch.Tensor, seg: torch.Tensor, sd: Dict[str, torch.Tensor], patch_size: int,
                          num_heads: int, window: int, num_superpixels: int,
                          img_size: Optional[int] = None) -> torch.Tensor:
    B = x.shape[0]
    img_size = img_size or x.shape[-1]
    emb = patch_embed(x, sd["patch_embed.projection.1.weight"], sd["patch_embed.projection.1.bias"], patch_size)
    pooled = []
    for b in range(B):
        d = map_patches_oracle(seg[b].numpy(), patch_size, img_size)
        pooled.append(pool_mean_oracle_autograd(emb[b], d))
    t = torch.stack(pooled)
    t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1)
    t = dynamic_positional_encoding(t, superpixel_centroids(seg, num_superpixels))
    for i in range(_depth(sd)):
        t = block(t, sd, f"blocks.{i}.", num_heads, window)
    D = t.shape[-1]
    t = F.layer_norm(t, (D,), sd["norm.weight"], sd["norm.bias"])
    return F.linear(t[:, 0], sd["head.weight"], sd["head.bias"])


def pool_mean_oracle_autograd(emb: torch.Tensor, d) -> torch.Tensor:
    """pool_mean_oracle written without in-place row writes into a leaf so that autograd flows (sppp.py:192-223)."""
    rows = [emb[patches, :].mean(dim=0) if patches else torch.zeros(emb.shape[-1]) for patches in d.values()]
    return torch.stack(rows).float()


__all__ = ["vit_mhla_forward", "sppp_vit_mhla_forward", "superpixel_centroids", "dynamic_positional_encoding",
           "patch_embed", "block", "pool_mean_oracle"]
