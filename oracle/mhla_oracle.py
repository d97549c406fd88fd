"""CPU oracle for Multi-Head Latent Attention.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates `/root/reference/models/mhla.py`:
  * window index table            -> mhla.py:46-83   (`window_indices`)
  * qkv projection + head split   -> mhla.py:100-102
  * latent projection of K and V  -> mhla.py:105-106
  * window gather, scaled scores  -> mhla.py:109-133
  * mask, softmax (dropout p=0)   -> mhla.py:136-147
  * PV, head merge, out proj      -> mhla.py:151-159

Two formulations are kept on purpose:
  `mhla_forward_gather`       follows the reference's own order of operations (index table,
                              two gathers, batched GEMVs).  It is the "port" that bench.py times
                              as the CPU baseline.
  `mhla_forward_closed_form`  the banded softmax with integer multiplicities and the latent
                              projection folded into the q / proj weights.  This is
                              the algebra the CUDA path implements; agreeing with the gather form
                              (and with the golden vectors of the real reference) pins it.

Parity pin: tests/golden/mhla_*.npz, produced by running the reference (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------
# integer part: the window table  (reference mhla.py:46-83)
# ----------------------------------------------------------------------------------------------
def window_row(i: int, seq_len: int, window_size: int) -> list:
    """One row of the reference's table.  mhla.py:63-81."""
    h = window_size // 2
    s = max(0, i - h)
    e = min(seq_len, i + h + 1)
    row = list(range(s, e))
    if len(row) < window_size:
        pad = window_size - len(row)
        if s == 0:
            row = row + [seq_len - 1] * pad      # mhla.py:74-76 (pad at the end with N-1)
        else:
            row = [0] * pad + row                # mhla.py:77-79 (pad at the front with 0)
    return row


def window_indices(seq_len: int, window_size: int) -> np.ndarray:
    """[N, W] int64 table; RuntimeError when rows are ragged (even W with N > W), like
    `torch.stack` at mhla.py:83."""
    rows = [window_row(i, seq_len, window_size) for i in range(seq_len)]
    lens = {len(r) for r in rows}
    if len(lens) > 1:
        raise RuntimeError(
            f"stack expects each tensor to be equal size (window_size={window_size}, seq_len={seq_len})")
    return np.asarray(rows, dtype=np.int64).reshape(seq_len, -1)


def window_multiplicity(seq_len: int, window_size: int) -> np.ndarray:
    """m[i, j] = how many of query i's window slots point at key j (SURVEY.md §8a closed form)."""
    m = np.zeros((seq_len, seq_len), dtype=np.int64)
    for i in range(seq_len):
        for j in window_row(i, seq_len, window_size):
            m[i, j] += 1
    return m


# ----------------------------------------------------------------------------------------------
# floating-point part, reference order of operations
# ----------------------------------------------------------------------------------------------
def mhla_forward_gather(
    x: torch.Tensor,
    qkv_w: torch.Tensor, qkv_b: torch.Tensor,
    proj_w: torch.Tensor, proj_b: torch.Tensor,
    lat_w: torch.Tensor, lat_b: torch.Tensor,
    num_heads: int, window_size: int,
    attention_mask: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Gather formulation, op for op as mhla.py:85-161 with dropout p = 0."""
    B, N, D = x.shape
    hd = D // num_heads
    qkv = torch.nn.functional.linear(x, qkv_w, qkv_b).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    k_lat = torch.nn.functional.linear(k, lat_w, lat_b)
    v_lat = torch.nn.functional.linear(v, lat_w, lat_b)
    idx = torch.from_numpy(window_indices(N, window_size)).to(x.device)          # [N, W]
    W = idx.shape[1]
    idx_b = idx[None, None].expand(B, num_heads, -1, -1)
    gidx = idx_b.unsqueeze(-1).expand(-1, -1, -1, -1, hd)
    k_win = torch.gather(k_lat.unsqueeze(3).expand(-1, -1, -1, W, -1), 2, gidx)
    v_win = torch.gather(v_lat.unsqueeze(3).expand(-1, -1, -1, W, -1), 2, gidx)
    attn = torch.matmul(q.unsqueeze(3), k_win.transpose(-2, -1)).squeeze(3) / (hd ** 0.5)
    if attention_mask is not None:
        wmask = torch.gather(attention_mask.unsqueeze(1).expand(-1, num_heads, -1, -1), 3, idx_b)
        attn = attn.masked_fill(wmask == 0, float("-inf"))
    attn = torch.softmax(attn, dim=-1)
    out = torch.matmul(attn.unsqueeze(3), v_win).squeeze(3)
    out = out.transpose(1, 2).reshape(B, N, D)
    return torch.nn.functional.linear(out, proj_w, proj_b)


# ----------------------------------------------------------------------------------------------
# attention-probability dropout (mhla.py:147): the keep-mask of favit's counter-based generator
# ----------------------------------------------------------------------------------------------
def dropout_keep_mask(B: int, H: int, N: int, W: int, p: float, seed: int) -> np.ndarray:
    """bool [B,H,N,W]: slot `pos` of row (b,h,i) is kept iff the splitmix64 finaliser of
    seed + (((b*H + h)*N + i)*W + pos) * 0x9E3779B97F4A7C15 gives a 24-bit uniform >= p.  The reference uses torch's
    Philox stream, so only the DISTRIBUTION is the reference's; this function pins the kernels' mask bit for bit."""
    c = np.arange(B * H * N * W, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + c * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return (u >= np.float32(p)).reshape(B, H, N, W)


def mlp_dropout_keep_mask(M: int, N: int, p: float, seed: int, layer: int, site: int):
    """(keep bool [M,N], inv_keep) of the fused MLP dropout (include/favit.h: favit_linear_fwd_dropout): element (row, col)
    belongs to group g = row * ceil(N/4) + col // 4; splitmix64(seed + offset(layer, site) + g * golden) gives four 16-bit
    uniforms; lane col % 4 is kept iff it is >= thr = round(p * 65536); kept values are scaled by 65536 / (65536 - thr).
    site 1 = after the activation, site 2 = after fc2 (models/vit.py:131-138).  The reference draws from torch's Philox
    stream: only the distribution is the reference's; this pins the kernels' mask bit for bit."""
    thr = int(np.rint(np.float32(p) * np.float32(65536.0)))
    gpr = (N + 3) // 4
    offset = ((2 * layer + site) * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
    g = (np.arange(M, dtype=np.uint64)[:, None] * np.uint64(gpr) + np.arange(gpr, dtype=np.uint64)[None, :])
    with np.errstate(over="ignore"):
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.uint64(offset) + g * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    lanes = np.stack([(z >> np.uint64(16 * t)) & np.uint64(0xFFFF) for t in range(4)], axis=-1).reshape(M, gpr * 4)[:, :N]
    return lanes >= np.uint64(thr), 65536.0 / (65536.0 - thr)


def mhla_attn_core_gather(q, k, v, window_size: int, attention_mask: Optional[torch.Tensor] = None,
                          keep: Optional[torch.Tensor] = None, dropout_p: float = 0.0) -> torch.Tensor:
    """The attention core in the reference's own formulation (mhla.py:109-154): window gather, scaled scores, mask,
    softmax over the W slots, dropout (`keep` [B,H,N,W] bool, scaled by 1/(1-p)), PV.  q,k,v [B,H,N,hd] -> [B,H,N,hd]."""
    B, H, N, hd = q.shape
    idx = torch.from_numpy(window_indices(N, window_size)).to(q.device)
    W = idx.shape[1]
    idx_b = idx[None, None].expand(B, H, -1, -1)
    gidx = idx_b.unsqueeze(-1).expand(-1, -1, -1, -1, hd)
    k_win = torch.gather(k.unsqueeze(3).expand(-1, -1, -1, W, -1), 2, gidx)
    v_win = torch.gather(v.unsqueeze(3).expand(-1, -1, -1, W, -1), 2, gidx)
    attn = torch.matmul(q.unsqueeze(3), k_win.transpose(-2, -1)).squeeze(3) / (hd ** 0.5)
    if attention_mask is not None:
        wmask = torch.gather(attention_mask.unsqueeze(1).expand(-1, H, -1, -1), 3, idx_b)
        attn = attn.masked_fill(wmask == 0, float("-inf"))
    attn = torch.softmax(attn, dim=-1)
    if keep is not None:
        attn = attn * keep.to(attn.dtype) / (1.0 - dropout_p)
    return torch.matmul(attn.unsqueeze(3), v_win).squeeze(3)


# ----------------------------------------------------------------------------------------------
# closed form: banded softmax with multiplicities + folded latent projection
# ----------------------------------------------------------------------------------------------
def fold_latent(qkv_w, qkv_b, proj_w, proj_b, lat_w, lat_b, num_heads: int):
    """Fold latent_proj (mhla.py:41,105-106) into the q rows of qkv and into proj.

    K path: q.(Wl k + bl) = (Wl^T q).k + q.bl; the second term is constant along the softmax axis
    and cancels.  V path: sum_j p_j (Wl v_j + bl) = Wl (sum_j p_j v_j) + bl because rows sum to 1.
    Row-vector convention (q = x Wq^T + bq):  Wq'_h = Wl^T Wq_h,  bq'_h = bq_h Wl,
    Wp' = Wp blockdiag_H(Wl),  bp' = Wp tile_H(bl) + bp.
    """
    D = proj_w.shape[0]
    hd = D // num_heads
    wq = qkv_w[:D].reshape(num_heads, hd, D)
    wq_f = torch.matmul(lat_w.t(), wq).reshape(D, D)
    bq_f = torch.matmul(qkv_b[:D].reshape(num_heads, hd), lat_w).reshape(D)
    qkv_w_f = torch.cat([wq_f, qkv_w[D:]], dim=0)
    qkv_b_f = torch.cat([bq_f, qkv_b[D:]], dim=0)
    wp = proj_w.reshape(D, num_heads, hd)
    proj_w_f = torch.matmul(wp, lat_w).reshape(D, D)
    proj_b_f = proj_b + torch.matmul(wp, lat_b).sum(dim=1)
    return qkv_w_f, qkv_b_f, proj_w_f, proj_b_f


def mhla_attn_core_closed_form(q, k, v, window_size: int,
                               attention_mask: Optional[torch.Tensor] = None
                               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """out_i = sum_j m_ij exp(s_ij) v_j / sum_j m_ij exp(s_ij); q,k,v [B,H,N,hd] -> (out, lse[B,H,N])."""
    B, H, N, hd = q.shape
    m = torch.from_numpy(window_multiplicity(N, window_size)).to(q.device)
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd)
    bias = torch.where(m > 0, torch.log(m.to(s.dtype)), torch.full_like(m, float("-inf"), dtype=s.dtype))
    s = s + bias
    if attention_mask is not None:
        s = s.masked_fill((attention_mask == 0).unsqueeze(1), float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, v), lse


def mhla_forward_closed_form(x, qkv_w, qkv_b, proj_w, proj_b, lat_w, lat_b,
                             num_heads: int, window_size: int,
                             attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, N, D = x.shape
    hd = D // num_heads
    if window_size % 2 == 0 and N > window_size:
        raise RuntimeError("stack expects each tensor to be equal size (even window_size)")
    qw, qb, pw, pb = fold_latent(qkv_w, qkv_b, proj_w, proj_b, lat_w, lat_b, num_heads)
    qkv = torch.nn.functional.linear(x, qw, qb).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    out, _ = mhla_attn_core_closed_form(qkv[0], qkv[1], qkv[2], window_size, attention_mask)
    out = out.transpose(1, 2).reshape(B, N, D)
    return torch.nn.functional.linear(out, pw, pb)
