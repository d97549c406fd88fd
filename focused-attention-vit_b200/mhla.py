"""Drop-in `MultiHeadLatentAttention` / `MHLATransformerBlock` backed by the favit sm_100a kernels.

Same constructors, attribute names (`qkv`, `proj`, `latent_proj`, `attn_dropout`, `proj_dropout`; block:
`norm1`, `attn`, `norm2`, `mlp`), state_dict keys and forward signatures as /root/reference/models/mhla.py:17-222,
so `vit_mhla.py`, `sppp_mhla.py` and `mhla_models.py` can swap these in.

What runs where:
  * the latent projection (mhla.py:105-106) is folded into the q rows of `qkv` and into `proj` with a handful of
    [hd x hd] torch matmuls per call (O(D^2 hd), independent of B*N); autograd through the fold gives the exact
    `latent_proj` gradients;
  * qkv / proj linears: `favit::linear` (tcgen05 GEMM in bf16, SIMT GEMM in fp32), fwd + dgrad + wgrad;
  * window gather + scores + mask + softmax + PV (mhla.py:109-154): `favit::mhla_attn`, forward and backward.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops


def compute_dtype(x: torch.Tensor) -> torch.dtype:
    """bf16 under CUDA autocast (custom ops are not autocast-aware, so the cast the reference gets from autocast is
    made explicitly), otherwise the input's own dtype."""
    if torch.is_autocast_enabled("cuda"):
        return torch.get_autocast_dtype("cuda")
    return x.dtype


def fold_latent(qkv_w, qkv_b, proj_w, proj_b, lat_w, lat_b, num_heads: int):
    """Wq'_h = Wl^T Wq_h, bq'_h = bq_h Wl, Wp' = Wp blockdiag_H(Wl), bp' = Wp tile_H(bl) + bp.

    K path: q.(Wl k + bl) = (Wl^T q).k + q.bl and q.bl is constant along the softmax axis; V path: softmax rows
    sum to one, so Wl and bl move behind the PV product and into proj."""
    D = proj_w.shape[0]
    hd = D // num_heads
    wq = torch.matmul(lat_w.t(), qkv_w[:D].reshape(num_heads, hd, D)).reshape(D, D)
    bq = torch.matmul(qkv_b[:D].reshape(num_heads, hd), lat_w).reshape(D)
    wp3 = proj_w.reshape(D, num_heads, hd)
    wp = torch.matmul(wp3, lat_w).reshape(D, D)
    bp = proj_b + torch.matmul(wp3, lat_b).sum(dim=1)
    return torch.cat([wq, qkv_w[D:]], dim=0), torch.cat([bq, qkv_b[D:]], dim=0), wp, bp


class MultiHeadLatentAttention(nn.Module):
    """Window-based "latent" attention (reference: models/mhla.py:17-161)."""

    def __init__(self, embed_dim: int, num_heads: int, window_size: int = 7, dropout: float = 0.0):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.window_size = window_size
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        self.qkv = nn.Linear(embed_dim, embed_dim * 3)
        self.proj = nn.Linear(embed_dim, embed_dim)
        self.latent_proj = nn.Linear(self.head_dim, self.head_dim)
        self.attn_dropout = nn.Dropout(dropout)
        self.proj_dropout = nn.Dropout(dropout)

    def _get_window_indices(self, seq_len: int) -> torch.Tensor:
        """[seq_len, window] int64 table of mhla.py:46-83 (kept for API compatibility; the kernels use the closed
        form and never read a table).  Raises RuntimeError for ragged rows (even window with seq_len > window),
        like `torch.stack` at mhla.py:83."""
        W, h, N = self.window_size, self.window_size // 2, seq_len
        i = torch.arange(N).unsqueeze(1)
        s = (i - h).clamp_min(0)
        e = (i + h + 1).clamp_max(N)
        if N > 0 and int((e - s).max()) > W:
            raise RuntimeError(f"stack expects each tensor to be equal size, but a window row has "
                               f"{int((e - s).max())} entries for window_size={W} (seq_len={N})")
        t = torch.arange(W).unsqueeze(0)
        pad = W - (e - s)
        left = s == 0                                    # pad at the end with N-1
        idx_left = torch.where(t < (e - s), s + t, torch.full_like(t + s, N - 1))
        idx_right = torch.where(t < pad, torch.zeros_like(t + s), s + t - pad)   # pad at the front with 0
        return torch.where(left, idx_left, idx_right).to(torch.int64)

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, N, D = x.shape
        W = self.window_size
        if W % 2 == 0 and N > W:
            raise RuntimeError(f"stack expects each tensor to be equal size: even window_size={W} with "
                               f"seq_len={N} > window_size is ragged (reference models/mhla.py:83)")
        cd = compute_dtype(x)
        self._check_supported(x, cd)
        if self.training and self.attn_dropout.p > 0:
            return self._forward_attn_dropout(x, attention_mask, cd)
        with torch.autocast("cuda", enabled=False):
            qw, qb, pw, pb = fold_latent(self.qkv.weight.float(), self.qkv.bias.float(), self.proj.weight.float(),
                                         self.proj.bias.float(), self.latent_proj.weight.float(),
                                         self.latent_proj.bias.float(), self.num_heads)
            xc = x if x.dtype == cd else x.to(cd)
            qkv = ops.linear(xc.reshape(B * N, D), qw, qb)
            mask = None
            if attention_mask is not None:
                mask = (attention_mask != 0).to(torch.uint8).contiguous()
            out, _ = ops.mhla_attn(qkv.view(B, N, 3, self.num_heads, self.head_dim), W, mask)
            y = ops.linear(out.reshape(B * N, D), pw, pb).view(B, N, D)
        return self.proj_dropout(y)

    def _check_supported(self, x: torch.Tensor, cd: torch.dtype) -> None:
        """The limits of the favit kernels, named up front (INTEGRATION.md lists them) instead of failing inside a kernel
        wrapper: the reference itself runs any dtype / head_dim on eager PyTorch."""
        if not x.is_cuda:
            raise RuntimeError("favit MultiHeadLatentAttention runs on CUDA tensors only (sm_100a kernels; there is no "
                               "CPU fallback)")
        if cd not in (torch.float32, torch.bfloat16):
            raise TypeError(f"favit MultiHeadLatentAttention computes in float32 or bfloat16 (bf16 autocast), got {cd}: "
                            "float16 / float64 inputs and fp16 autocast are not supported")
        if self.head_dim not in (16, 32, 64, 128):
            raise ValueError(f"favit MultiHeadLatentAttention supports head_dim 16, 32, 64 or 128, got {self.head_dim} "
                             f"(embed_dim {self.embed_dim} / num_heads {self.num_heads})")
        if cd == torch.bfloat16 and self.embed_dim % 8:
            raise ValueError(f"favit MultiHeadLatentAttention needs embed_dim % 8 == 0 in bf16 (TMA row pitch), got "
                             f"{self.embed_dim}")

    def _forward_attn_dropout(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor], cd) -> torch.Tensor:
        """Training with attention-probability dropout p > 0 (mhla.py:147).  The K-side fold of latent_proj stays valid
        (its bias term is constant along the softmax axis and dropout comes after the softmax); the V-side fold does
        not, because dropped rows of probabilities no longer sum to one, so V gets its latent projection explicitly as
        in mhla.py:106 and proj keeps its own weights.  Statistical parity with the reference: the keep-mask comes from
        favit's counter-based generator, one Bernoulli per window slot like nn.Dropout on the [B,H,N,W] probabilities."""
        if torch.cuda.is_current_stream_capturing():
            # the keep-mask seed of the attention kernels is a host value baked into the launch arguments: a captured
            # graph would replay the same mask on every step (nn.Dropout stays random under capture)
            raise RuntimeError("attention-probability dropout (attn_dropout > 0 in training mode) cannot be captured in a "
                               "CUDA graph: use engine.TrainStep(cuda_graph=False) or attn_dropout=0.0")
        B, N, D = x.shape
        H, hd, W = self.num_heads, self.head_dim, self.window_size
        p = float(self.attn_dropout.p)
        with torch.autocast("cuda", enabled=False):
            qkv_w, qkv_b = self.qkv.weight.float(), self.qkv.bias.float()
            lat_w, lat_b = self.latent_proj.weight.float(), self.latent_proj.bias.float()
            wq = torch.matmul(lat_w.t(), qkv_w[:D].reshape(H, hd, D)).reshape(D, D)       # Wq'_h = Wl^T Wq_h
            bq = torch.matmul(qkv_b[:D].reshape(H, hd), lat_w).reshape(D)                  # bq'_h = bq_h Wl
            xc = x if x.dtype == cd else x.to(cd)
            qkv = ops.linear(xc.reshape(B * N, D), torch.cat([wq, qkv_w[D:]], dim=0), torch.cat([bq, qkv_b[D:]], dim=0))
            qkv = qkv.view(B, N, 3, H, hd)
            v_lat = ops.linear(qkv[:, :, 2].reshape(B * N * H, hd), lat_w, lat_b).view(B, N, H, hd)
            packed = torch.stack([qkv[:, :, 0], qkv[:, :, 1], v_lat], dim=2).contiguous()
            mask = None
            if attention_mask is not None:
                mask = (attention_mask != 0).to(torch.uint8).contiguous()
            out, _ = ops.mhla_attn(packed, W, mask, p, _dropout_seed())
            y = ops.linear(out.reshape(B * N, D), self.proj.weight.float(), self.proj.bias.float()).view(B, N, D)
        return self.proj_dropout(y)


def _dropout_seed() -> int:
    """One 62-bit seed per forward from torch's CPU generator (reproducible under torch.manual_seed).  It is a host
    value: `_forward_attn_dropout` refuses to run under CUDA-graph capture, where it would be frozen."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


class MHLATransformerBlock(nn.Module):
    """Transformer block with MHLA (reference: models/mhla.py:164-222)."""

    def __init__(self, embed_dim: int, num_heads: int, window_size: int = 7, mlp_ratio: float = 4.0,
                 dropout: float = 0.0, attn_dropout: float = 0.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(embed_dim)
        self.attn = MultiHeadLatentAttention(embed_dim=embed_dim, num_heads=num_heads, window_size=window_size,
                                             dropout=attn_dropout)
        self.norm2 = nn.LayerNorm(embed_dim)
        mlp_hidden_dim = int(embed_dim * mlp_ratio)
        self.mlp = nn.Sequential(
            nn.Linear(embed_dim, mlp_hidden_dim),
            nn.GELU(),
            nn.Dropout(dropout),
            nn.Linear(mlp_hidden_dim, embed_dim),
            nn.Dropout(dropout),
        )

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        from . import fused_block
        cd = compute_dtype(x)
        if fused_block.fusable(x, self.attn, self.mlp[2].p, self.training, attention_mask, cd,
                               self.mlp[0].out_features):
            with torch.autocast("cuda", enabled=False):
                return fused_block.fused_block(x, self.norm1, self.attn, self.norm2, self.mlp[0], self.mlp[3], cd,
                                               self.mlp[2].p if self.training else 0.0)
        x = x + self.attn(self.norm1(x), attention_mask)
        x = x + self.mlp(self.norm2(x))
        return x
