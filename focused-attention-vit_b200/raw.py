"""Thin tensor-level wrappers over the C-ABI (no dispatcher, no autograd): the building blocks of the fused
transformer-block ops in `fused_block.py`.  All tensors are CUDA, contiguous and 2-D unless stated."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


# layers / dropout sites of one forward call draw from the same device seed at different offsets
_DROP_STRIDE = 0xD1B54A32D192ED03


_TILE_MODES = {"static": 0, "steal": 1}


def gemm_tile_scheduler(mode: Optional[str] = None) -> str:
    """How the persistent CTA-pair GEMM hands out its tiles (include/favit.h: favit_set_gemm_tile_scheduler): "static"
    striding, or "steal" = work stealing for launches that share the GPU with another kernel (the overlapped NCCL
    all-reduce).  Process-wide, applies to launches / graph captures made afterwards; None only queries.  Returns the
    mode in force."""
    if mode is not None and mode not in _TILE_MODES:
        raise ValueError(f"unknown tile scheduler {mode!r}: 'static' or 'steal'")
    got = L.lib().favit_set_gemm_tile_scheduler(-1 if mode is None else _TILE_MODES[mode])
    return "steal" if got == 1 else "static"


def drop_offset(layer: int, site: int) -> int:
    """site 1 = the dropout after the activation, site 2 = the dropout after fc2 (models/vit.py:131-138)."""
    return ((2 * layer + site) * _DROP_STRIDE) & 0xFFFFFFFFFFFFFFFF


def linear_fwd(x: Tensor, w: Tensor, bias: Optional[Tensor], residual: Optional[Tensor], out_dtype: torch.dtype,
               gelu: bool = False, save_preact: bool = False, drop=None) -> Tuple[Tensor, Optional[Tensor]]:
    """drop = (p, seed int64[1] device tensor, offset): dropout fused behind the activation (favit_linear_fwd_dropout)."""
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=out_dtype, device=x.device)
    pre = torch.empty((M, N), dtype=x.dtype, device=x.device) if (gelu and save_preact) else None
    p_, seed_, off_ = drop if drop is not None else (0.0, None, 0)
    rc = L.call("gemm_fwd", 2.0 * M * N * K, L.lib().favit_linear_fwd_dropout, _p(x), _p(w), _p(bias), _p(residual), _p(y),
                _p(pre), M, N, K, K, K, N, N, _DT[x.dtype], _DT[out_dtype],
                _DT[residual.dtype] if residual is not None else L.F32, L.EPI_GELU if gelu else L.EPI_NONE,
                float(p_), _p(seed_), off_, _s())
    L.check(rc, "favit_linear_fwd")
    return y, pre


def dropout_cast(g: Tensor, out_dtype: torch.dtype, drop, colsum: Optional[Tensor] = None) -> Tensor:
    """keep / (1 - p) * g (fp32 [M,N]) in out_dtype; `colsum` (already zeroed fp32 [N]) receives the column sums."""
    M, N = g.shape
    out = torch.empty((M, N), dtype=out_dtype, device=g.device)
    p_, seed_, off_ = drop
    rc = L.call("dropout_cast", float(g.numel() * (4 + out.element_size())), L.lib().favit_dropout_cast, _p(g), _p(out),
                _DT[out_dtype], _p(colsum), M, N, float(p_), _p(seed_), off_, _s())
    L.check(rc, "favit_dropout_cast")
    return out


def linear_dgrad(dy: Tensor, w: Tensor, preact: Optional[Tensor], out_dtype: torch.dtype, colsum: bool = False,
                 zeroed: Optional[Tensor] = None, drop=None):
    """dx = dy @ w (* gelu'(preact)); with colsum also the fp32 column sums of dx (returns (dx, sums)).
    `zeroed`: an already zeroed fp32 [K] buffer to accumulate the column sums into (saves a fill launch)."""
    M, N = dy.shape
    K = w.shape[1]
    dx = torch.empty((M, K), dtype=out_dtype, device=dy.device)
    sums = (zeroed if zeroed is not None else torch.zeros((K,), dtype=torch.float32, device=dy.device)) if colsum else None
    p_, seed_, off_ = drop if drop is not None else (0.0, None, 0)
    rc = L.call("gemm_dgrad", 2.0 * M * N * K, L.lib().favit_linear_dgrad_dropout, _p(dy), _p(w), _p(preact), _p(dx),
                _p(sums), M, N, K, N, K, K, _DT[dy.dtype], _DT[out_dtype],
                L.EPI_DGELU_MUL if preact is not None else L.EPI_NONE, float(p_), _p(seed_), off_, _s())
    L.check(rc, "favit_linear_dgrad")
    return (dx, sums) if colsum else dx


def linear_wgrad(dy: Tensor, x: Tensor, want_bias: bool = True) -> Tuple[Tensor, Optional[Tensor]]:
    M, N = dy.shape
    K = x.shape[1]
    dw = torch.empty((N, K), dtype=torch.float32, device=dy.device)
    db = torch.empty((N,), dtype=torch.float32, device=dy.device) if want_bias else None
    rc = L.call("gemm_wgrad", 2.0 * M * N * K, L.lib().favit_linear_wgrad, _p(dy), _p(x), _p(dw), _p(db), M, N, K, N, K,
                K, _DT[dy.dtype], 0, _s())
    L.check(rc, "favit_linear_wgrad")
    return dw, db


def ln_fwd(x: Tensor, gamma: Tensor, beta: Tensor, out_dtype: torch.dtype, eps: float,
           delta: Optional[Tensor] = None):
    """Returns (y, mean, rstd) or, with delta (out_dtype), (y, mean, rstd, xsum = x + delta)."""
    M, D = x.shape
    y = torch.empty((M, D), dtype=out_dtype, device=x.device)
    xsum = torch.empty_like(x) if delta is not None else None
    mean = torch.empty((M,), dtype=torch.float32, device=x.device)
    rstd = torch.empty((M,), dtype=torch.float32, device=x.device)
    work = float(x.numel() * x.element_size() + y.numel() * y.element_size())
    if delta is not None:
        work += float(delta.numel() * delta.element_size() + xsum.numel() * xsum.element_size())
    rc = L.call("ln_fwd", work, L.lib().favit_layernorm_fwd, _p(x), _DT[x.dtype], _p(delta), _p(xsum), _p(gamma),
                _p(beta), _p(y), _DT[out_dtype], _p(mean), _p(rstd), M, D, float(eps), _s())
    L.check(rc, "favit_layernorm_fwd")
    if delta is not None:
        return y, mean, rstd, xsum
    return y, mean, rstd


def ln_bwd(dy: Tensor, x: Tensor, mean: Tensor, rstd: Tensor, gamma: Tensor, dres: Optional[Tensor],
           want_bf16: bool, zeroed: Optional[Tensor] = None):
    """Returns (dx fp32, dx_bf16 or None, dgamma, dbeta, column sums of dx).  `zeroed`: an already zeroed fp32 [3, D]
    buffer for the three accumulators (the results are then views of it, not copies)."""
    M, D = x.shape
    dx = torch.empty((M, D), dtype=torch.float32, device=x.device)
    dxb = torch.empty((M, D), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    acc = zeroed if zeroed is not None else torch.zeros((3, D), dtype=torch.float32, device=x.device)
    dg, db, dxs = acc[0], acc[1], acc[2]
    work = float(dy.numel() * dy.element_size() + x.numel() * x.element_size() + dx.numel() * 4 +
                 (dres.numel() * 4 if dres is not None else 0) + (dxb.numel() * 2 if want_bf16 else 0))
    rc = L.call("ln_bwd", work, L.lib().favit_layernorm_bwd, _p(dy), _DT[dy.dtype], _p(x), _DT[x.dtype], _p(mean),
                _p(rstd), _p(gamma), _p(dres), _p(dx), _p(dxb), dg.data_ptr(), db.data_ptr(), dxs.data_ptr(), M, D, _s())
    L.check(rc, "favit_layernorm_bwd")
    if zeroed is not None:
        return dx, dxb, dg, db, dxs
    return dx, dxb, dg.clone(), db.clone(), dxs.clone()


def attn_fwd(qkv: Tensor, B: int, N: int, H: int, hd: int, window: int):
    """qkv [B*N, 3*H*hd] (packed, contiguous) -> out [B*N, H*hd], lse [B,H,N]."""
    es = qkv.element_size()
    out = torch.empty((B * N, H * hd), dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty((B, H, N), dtype=torch.float32, device=qkv.device)
    base = qkv.data_ptr()
    rc = L.call("attn_fwd", 4.0 * B * N * H * hd * es, L.lib().favit_mhla_attn_fwd, base, base + H * hd * es,
                base + 2 * H * hd * es, None, _p(out), _p(lse), B, H, N, hd, window, float(hd) ** -0.5,
                N * 3 * H * hd, 3 * H * hd, hd, _DT[qkv.dtype], 0.0, 0, _s())
    L.check(rc, "favit_mhla_attn_fwd")
    return out, lse


def attn_bwd(qkv: Tensor, out: Tensor, lse: Tensor, dout: Tensor, B: int, N: int, H: int, hd: int, window: int,
             zeroed: Optional[Tensor] = None):
    """Returns (dqkv, fp32 column sums of dqkv = the qkv bias gradient).  `zeroed`: an already zeroed fp32 [3*H*hd]."""
    es = qkv.element_size()
    dqkv = torch.empty_like(qkv)
    sums = zeroed if zeroed is not None else torch.zeros((3 * H * hd,), dtype=torch.float32, device=qkv.device)
    delta = torch.empty((B, H, N), dtype=torch.float32, device=qkv.device)
    base, dbase = qkv.data_ptr(), dqkv.data_ptr()
    off = H * hd * es
    rc = L.call("attn_bwd", 8.0 * B * N * H * hd * es, L.lib().favit_mhla_attn_bwd, base, base + off, base + 2 * off,
                None, _p(out), _p(lse), _p(dout), dbase, dbase + off, dbase + 2 * off, _p(delta), _p(sums), B, H, N, hd,
                window, float(hd) ** -0.5, N * 3 * H * hd, 3 * H * hd, hd, _DT[qkv.dtype], 0.0, 0, _s())
    L.check(rc, "favit_mhla_attn_bwd")
    return dqkv, sums


def fold_fwd(qkv_w: Tensor, qkv_b: Tensor, proj_w: Tensor, proj_b: Tensor, lat_w: Tensor, lat_b: Tensor, H: int,
             cd: torch.dtype):
    """fp32 masters -> (wqkv' [3D,D] cd, bqkv' [3D] fp32, wproj' [D,D] cd, bproj' [D] fp32) with latent_proj folded in."""
    D = proj_w.shape[0]
    hd = D // H
    dev = qkv_w.device
    wq = torch.empty((3 * D, D), dtype=cd, device=dev)
    wp = torch.empty((D, D), dtype=cd, device=dev)
    bq = torch.empty((3 * D,), dtype=torch.float32, device=dev)
    bp = torch.empty((D,), dtype=torch.float32, device=dev)
    rc = L.call("fold", 0.0, L.lib().favit_latent_fold_fwd, _p(qkv_w), _p(qkv_b), _p(proj_w), _p(proj_b), _p(lat_w),
                _p(lat_b), _p(wq), _p(bq), _p(wp), _p(bp), H, hd, _DT[cd], _s())
    L.check(rc, "favit_latent_fold_fwd")
    return wq, bq, wp, bp


def cast_bf16_batched(tensors):
    """fp32 tensors -> bf16 copies, one launch for all of them (views of one flat buffer)."""
    import ctypes as C
    n = len(tensors)
    srcs = [t.detach().float().contiguous() for t in tensors]
    sizes = [t.numel() for t in srcs]
    pad = [(s + 7) // 8 * 8 for s in sizes]                      # keep every view 16-byte aligned
    flat = torch.empty((sum(pad),), dtype=torch.bfloat16, device=srcs[0].device)
    outs, off = [], 0
    for t, s, p in zip(srcs, sizes, pad):
        outs.append(flat[off:off + s].view(t.shape))
        off += p
    rc = L.call("cast", 0.0, L.lib().favit_cast_bf16_batched, n, (C.c_void_p * n)(*[t.data_ptr() for t in srcs]),
                (C.c_void_p * n)(*[o.data_ptr() for o in outs]), (C.c_int64 * n)(*sizes), _s())
    L.check(rc, "favit_cast_bf16_batched")
    return outs


def copy_batched(srcs, dsts) -> None:
    """dsts[i][...] = srcs[i] (fp32 / bf16, converted if the dtypes differ; all srcs one dtype, all dsts one dtype),
    one launch per 32 tensors.  Tensors must be contiguous."""
    import ctypes as C
    n = len(srcs)
    if n == 0:
        return
    rc = L.call("copy", 0.0, L.lib().favit_copy_batched, n, (C.c_void_p * n)(*[t.data_ptr() for t in srcs]),
                (C.c_void_p * n)(*[t.data_ptr() for t in dsts]), (C.c_int64 * n)(*[t.numel() for t in srcs]),
                _DT[srcs[0].dtype], _DT[dsts[0].dtype], _s())
    L.check(rc, "favit_copy_batched")


def fold_fwd_batched(layers, H: int, cd: torch.dtype):
    """layers: list of (qkv_w, qkv_b, proj_w, proj_b, lat_w, lat_b) fp32 tensors of the L blocks of one model.
    One call for the whole model; returns L tuples (wqkv' cd, bqkv' fp32, wproj' cd, bproj' fp32), views of four
    contiguous [L, ...] buffers."""
    import ctypes as C
    nl = len(layers)
    D = layers[0][2].shape[0]
    hd = D // H
    dev = layers[0][0].device
    wq = torch.empty((nl, 3 * D, D), dtype=cd, device=dev)
    wp = torch.empty((nl, D, D), dtype=cd, device=dev)
    bq = torch.empty((nl, 3 * D), dtype=torch.float32, device=dev)
    bp = torch.empty((nl, D), dtype=torch.float32, device=dev)
    ptrs = []
    for i, lay in enumerate(layers):
        ptrs += [t.data_ptr() for t in lay] + [wq[i].data_ptr(), bq[i].data_ptr(), wp[i].data_ptr(), bp[i].data_ptr()]
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    rc = L.call("fold", 0.0, L.lib().favit_latent_fold_fwd_batched, nl, arr, H, hd, _DT[cd], _s())
    L.check(rc, "favit_latent_fold_fwd_batched")
    return [(wq[i], bq[i], wp[i], bp[i]) for i in range(nl)]


def fold_bwd_batched(layers, grads, H: int):
    """layers: list of (qkv_w, qkv_b, proj_w, lat_w, lat_b); grads: list of (dwqkv, dbqkv, dwproj, dbproj), the
    gradients of the folded weights, rewritten IN PLACE into the gradients of qkv.weight / qkv.bias / proj.weight.
    Returns L tuples (dlat_w, dlat_b)."""
    import ctypes as C
    nl = len(layers)
    D = layers[0][2].shape[0]
    hd = D // H
    dl = torch.empty((nl, hd * hd + hd), dtype=torch.float32, device=layers[0][0].device)
    ptrs = []
    for i in range(nl):
        ptrs += [t.data_ptr() for t in layers[i]] + [t.data_ptr() for t in grads[i]]
        ptrs += [dl[i].data_ptr(), dl[i, hd * hd:].data_ptr()]
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    rc = L.call("fold", 0.0, L.lib().favit_latent_fold_bwd_batched, nl, arr, H, hd, _s())
    L.check(rc, "favit_latent_fold_bwd_batched")
    return [(dl[i, :hd * hd].view(hd, hd), dl[i, hd * hd:]) for i in range(nl)]


def fold_bwd(qkv_w: Tensor, qkv_b: Tensor, proj_w: Tensor, lat_w: Tensor, lat_b: Tensor, dwqkv: Tensor, dbqkv: Tensor,
             dwproj: Tensor, dbproj: Tensor, H: int):
    """In place: dwqkv / dbqkv / dwproj become the gradients of qkv.weight / qkv.bias / proj.weight.
    Returns (dlat_w, dlat_b)."""
    D = proj_w.shape[0]
    hd = D // H
    dlw = torch.empty((hd, hd), dtype=torch.float32, device=qkv_w.device)
    dlb = torch.empty((hd,), dtype=torch.float32, device=qkv_w.device)
    rc = L.call("fold", 0.0, L.lib().favit_latent_fold_bwd, _p(qkv_w), _p(qkv_b), _p(proj_w), _p(lat_w), _p(lat_b),
                _p(dwqkv), _p(dbqkv), _p(dwproj), _p(dbproj), _p(dlw), _p(dlb), H, hd, _s())
    L.check(rc, "favit_latent_fold_bwd")
    return dlw, dlb
