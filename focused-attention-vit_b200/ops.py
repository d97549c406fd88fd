"""torch.library registration of the favit C-ABI kernels (namespace ``favit::``).

Every op allocates its outputs with torch on the input's device, passes raw device pointers and the current
torch stream to libfavit_b200.so, and never touches the CPU oracle.  Differentiable ops (`linear`,
`mhla_attn`, `sppp_pool`) carry their backward through `register_autograd`, itself made of favit ops.

Reference spans replaced (under /root/reference): models/mhla.py:100,158 (linears), :109-154 (window attention),
models/sppp.py:91-128 (assignment), :192-223 + models/sppp_mhla.py:286-300 (pooling).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}
_TD = {L.F32: torch.float32, L.BF16: torch.bfloat16}


def _dt(t: Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"favit ops take float32 or bfloat16 tensors, got {t.dtype}") from None


def _cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("favit ops run on CUDA tensors only (sm_100a kernels; there is no CPU fallback)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _rowmajor(t: Tensor) -> Tensor:
    """2-D view whose last dim is contiguous (leading dimension = stride(0))."""
    if t.dim() != 2:
        raise ValueError("expected a 2-D tensor")
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def _ld(t: Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


# ------------------------------------------------------------------------------------------------
# linear layers
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("favit::linear_fwd", mutates_args=())
def linear_fwd(x: Tensor, w: Tensor, bias: Optional[Tensor], residual: Optional[Tensor], gelu: bool,
               out_dtype: torch.dtype, save_preact: bool) -> Tuple[Tensor, Tensor]:
    """y = act(x @ w.T + bias) + residual.  x [M,K], w [N,K] (same dtype), bias fp32 [N].  Returns (y, preact);
    preact is empty unless gelu and save_preact."""
    _cuda(x, w, bias, residual)
    x, w = _rowmajor(x), _rowmajor(w)
    M, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K or w.dtype != x.dtype:
        raise ValueError(f"linear_fwd: x {tuple(x.shape)} {x.dtype} vs w {tuple(w.shape)} {w.dtype}")
    if bias is not None:
        bias = bias.float().contiguous()
    y = torch.empty((M, N), dtype=out_dtype, device=x.device)
    pre = torch.empty((M, N) if (gelu and save_preact) else (0,), dtype=x.dtype, device=x.device)
    if residual is not None:
        residual = _rowmajor(residual)
    if M == 0:
        return y, pre
    rc = L.call("gemm_fwd", 2.0 * M * N * K, L.lib().favit_linear_fwd,
        _p(x), _p(w), _p(bias), _p(residual), _p(y), _p(pre) if pre.numel() else None, M, N, K, _ld(x), _ld(w), N,
        _ld(residual) if residual is not None else 0, _dt(x), _DT[out_dtype],
        _dt(residual) if residual is not None else L.F32, L.EPI_GELU if gelu else L.EPI_NONE, _stream())
    L.check(rc, "favit_linear_fwd")
    return y, pre


@linear_fwd.register_fake
def _(x, w, bias, residual, gelu, out_dtype, save_preact):
    M, N = x.shape[0], w.shape[0]
    return (x.new_empty((M, N), dtype=out_dtype),
            x.new_empty((M, N) if (gelu and save_preact) else (0,)))


@torch.library.custom_op("favit::linear_dgrad", mutates_args=())
def linear_dgrad(dy: Tensor, w: Tensor, preact: Optional[Tensor], out_dtype: torch.dtype) -> Tensor:
    """dx = dy @ w  (* gelu'(preact)).  dy [M,N], w [N,K]."""
    _cuda(dy, w, preact)
    dy, w = _rowmajor(dy), _rowmajor(w)
    M, N = dy.shape
    K = w.shape[1]
    if w.shape[0] != N or w.dtype != dy.dtype:
        raise ValueError(f"linear_dgrad: dy {tuple(dy.shape)} {dy.dtype} vs w {tuple(w.shape)} {w.dtype}")
    dx = torch.empty((M, K), dtype=out_dtype, device=dy.device)
    if preact is not None:
        preact = preact.contiguous()
    if M == 0:
        return dx
    rc = L.call("gemm_dgrad", 2.0 * M * N * K, L.lib().favit_linear_dgrad, _p(dy), _p(w), _p(preact), _p(dx), None, M, N, K, _ld(dy), _ld(w), K, _dt(dy),
                                    _DT[out_dtype], L.EPI_DGELU_MUL if preact is not None else L.EPI_NONE, _stream())
    L.check(rc, "favit_linear_dgrad")
    return dx


@linear_dgrad.register_fake
def _(dy, w, preact, out_dtype):
    return dy.new_empty((dy.shape[0], w.shape[1]), dtype=out_dtype)


@torch.library.custom_op("favit::linear_wgrad", mutates_args=())
def linear_wgrad(dy: Tensor, x: Tensor, want_bias: bool) -> Tuple[Tensor, Tensor]:
    """dw[N,K] = dy.T @ x (fp32), db[N] = column sums of dy (fp32; empty unless want_bias)."""
    _cuda(dy, x)
    dy, x = _rowmajor(dy), _rowmajor(x)
    M, N = dy.shape
    K = x.shape[1]
    if x.shape[0] != M or x.dtype != dy.dtype:
        raise ValueError(f"linear_wgrad: dy {tuple(dy.shape)} {dy.dtype} vs x {tuple(x.shape)} {x.dtype}")
    dw = torch.empty((N, K), dtype=torch.float32, device=dy.device)
    db = torch.empty((N,) if want_bias else (0,), dtype=torch.float32, device=dy.device)
    if M == 0:
        return dw.zero_(), db.zero_()
    rc = L.call("gemm_wgrad", 2.0 * M * N * K, L.lib().favit_linear_wgrad, _p(dy), _p(x), _p(dw), _p(db) if want_bias else None, M, N, K, _ld(dy), _ld(x), K,
                                    _dt(dy), 0, _stream())
    L.check(rc, "favit_linear_wgrad")
    return dw, db


@linear_wgrad.register_fake
def _(dy, x, want_bias):
    N, K = dy.shape[1], x.shape[1]
    return (dy.new_empty((N, K), dtype=torch.float32), dy.new_empty((N,) if want_bias else (0,), dtype=torch.float32))


@torch.library.custom_op("favit::linear", mutates_args=())
def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """nn.Linear forward in x's dtype: x [..., K] (fp32 or bf16), weight [N,K] / bias [N] of any float dtype
    (fp32 master parameters are cast to the compute dtype here; their gradients come back in fp32)."""
    K = x.shape[-1]
    x2 = x.reshape(-1, K)
    w = weight if weight.dtype == x.dtype else weight.to(x.dtype)
    y, _ = linear_fwd(x2, w, bias, None, False, x.dtype, False)
    return y.view(*x.shape[:-1], weight.shape[0])


@linear.register_fake
def _(x, weight, bias):
    return x.new_empty((*x.shape[:-1], weight.shape[0]))


def _linear_setup(ctx, inputs, output):
    x, weight, bias = inputs
    ctx.save_for_backward(x, weight)
    ctx.has_bias = bias is not None
    ctx.bias_dtype = bias.dtype if bias is not None else None


def _linear_backward(ctx, dy):
    x, weight = ctx.saved_tensors
    K = x.shape[-1]
    N = weight.shape[0]
    dy2 = dy.reshape(-1, N)
    if dy2.dtype != x.dtype:
        dy2 = dy2.to(x.dtype)
    x2 = x.reshape(-1, K)
    dx = dw = db = None
    if ctx.needs_input_grad[0]:
        w = weight if weight.dtype == x.dtype else weight.to(x.dtype)
        dx = linear_dgrad(dy2, w, None, x.dtype).view(x.shape)
    if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
        dw, db = linear_wgrad(dy2, x2, ctx.has_bias)
        dw = dw.to(weight.dtype)
        db = db.to(ctx.bias_dtype) if ctx.has_bias else None
    return dx, dw, db


linear.register_autograd(_linear_backward, setup_context=_linear_setup)


# ------------------------------------------------------------------------------------------------
# MHLA window attention core
# ------------------------------------------------------------------------------------------------
def _qkv_args(qkv: Tensor):
    B, N, three, H, hd = qkv.shape
    if three != 3 or not qkv.is_contiguous():
        raise ValueError("mhla_attn: qkv must be a contiguous [B, N, 3, H, hd] tensor (the packed qkv GEMM output)")
    es = qkv.element_size()
    base = qkv.data_ptr()
    return B, N, H, hd, base, base + H * hd * es, base + 2 * H * hd * es, N * 3 * H * hd, 3 * H * hd, hd


@torch.library.custom_op("favit::mhla_attn_fwd", mutates_args=())
def mhla_attn_fwd(qkv: Tensor, window: int, mask: Optional[Tensor], dropout_p: float = 0.0,
                  seed: int = 0) -> Tuple[Tensor, Tensor]:
    """qkv [B,N,3,H,hd] -> (out [B,N,H*hd], lse fp32 [B,H,N]).  mask: uint8 [B,N,N] or None.  dropout_p > 0:
    attention-probability dropout (mhla.py:147) with the counter-based mask of (seed, b, h, i, window slot)."""
    _cuda(qkv, mask)
    B, N, H, hd, q, k, v, sb, sn, sh = _qkv_args(qkv)
    out = torch.empty((B, N, H * hd), dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty((B, H, N), dtype=torch.float32, device=qkv.device)
    if mask is not None and (mask.dtype != torch.uint8 or mask.shape != (B, N, N) or not mask.is_contiguous()):
        raise ValueError("mhla_attn: mask must be a contiguous uint8 [B,N,N] tensor")
    if B * N == 0:
        return out, lse
    rc = L.call("attn_fwd", 4.0 * B * N * H * hd * qkv.element_size(), L.lib().favit_mhla_attn_fwd, q, k, v, _p(mask), _p(out), _p(lse), B, H, N, hd, window, float(hd) ** -0.5,
                                     sb, sn, sh, _dt(qkv), float(dropout_p), int(seed), _stream())
    L.check(rc, "favit_mhla_attn_fwd")
    return out, lse


@mhla_attn_fwd.register_fake
def _(qkv, window, mask, dropout_p=0.0, seed=0):
    B, N, _, H, hd = qkv.shape
    return qkv.new_empty((B, N, H * hd)), qkv.new_empty((B, H, N), dtype=torch.float32)


@torch.library.custom_op("favit::mhla_attn_bwd", mutates_args=())
def mhla_attn_bwd(qkv: Tensor, out: Tensor, lse: Tensor, dout: Tensor, window: int, mask: Optional[Tensor],
                  dropout_p: float = 0.0, seed: int = 0) -> Tensor:
    """Gradient of mhla_attn_fwd w.r.t. the packed qkv: returns dqkv [B,N,3,H,hd]."""
    _cuda(qkv, out, lse, dout, mask)
    B, N, H, hd, q, k, v, sb, sn, sh = _qkv_args(qkv)
    dout = dout.contiguous()
    if dout.dtype != qkv.dtype:
        dout = dout.to(qkv.dtype)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty((B, H, N), dtype=torch.float32, device=qkv.device)
    if B * N == 0:
        return dqkv
    es = qkv.element_size()
    dq = dqkv.data_ptr()
    rc = L.call("attn_bwd", 8.0 * B * N * H * hd * es, L.lib().favit_mhla_attn_bwd, q, k, v, _p(mask), _p(out), _p(lse), _p(dout), dq, dq + H * hd * es,
                                     dq + 2 * H * hd * es, _p(delta), None, B, H, N, hd, window, float(hd) ** -0.5,
                                     sb, sn, sh, _dt(qkv), float(dropout_p), int(seed), _stream())
    L.check(rc, "favit_mhla_attn_bwd")
    return dqkv


@mhla_attn_bwd.register_fake
def _(qkv, out, lse, dout, window, mask, dropout_p=0.0, seed=0):
    return torch.empty_like(qkv)


@torch.library.custom_op("favit::mhla_attn", mutates_args=())
def mhla_attn(qkv: Tensor, window: int, mask: Optional[Tensor], dropout_p: float = 0.0,
              seed: int = 0) -> Tuple[Tensor, Tensor]:
    """Differentiable window attention: (out [B,N,D], lse)."""
    return mhla_attn_fwd(qkv, window, mask, dropout_p, seed)


@mhla_attn.register_fake
def _(qkv, window, mask, dropout_p=0.0, seed=0):
    B, N, _, H, hd = qkv.shape
    return qkv.new_empty((B, N, H * hd)), qkv.new_empty((B, H, N), dtype=torch.float32)


def _attn_setup(ctx, inputs, output):
    qkv, window, mask, dropout_p, seed = inputs
    out, lse = output
    ctx.dropout_p, ctx.seed = dropout_p, seed
    ctx.save_for_backward(qkv, out, lse, mask if mask is not None else torch.empty(0))
    ctx.window = window
    ctx.has_mask = mask is not None
    ctx.set_materialize_grads(False)


def _attn_backward(ctx, dout, dlse):
    qkv, out, lse, mask = ctx.saved_tensors
    if dlse is not None:
        raise RuntimeError("mhla_attn: gradients through the log-sum-exp output are not supported")
    if dout is None:
        return torch.zeros_like(qkv), None, None, None, None
    return (mhla_attn_bwd(qkv, out, lse, dout, ctx.window, mask if ctx.has_mask else None, ctx.dropout_p, ctx.seed),
            None, None, None, None)


mhla_attn.register_autograd(_attn_backward, setup_context=_attn_setup)


# ------------------------------------------------------------------------------------------------
# SPPP
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("favit::sppp_assign", mutates_args=())
def sppp_assign(labels: Tensor, patch_size: int, img_size: int, r_cap: int) -> List[Tensor]:
    """labels int64 [B,H,W] -> [dom i64 [B,P], slot i32 [B,P], num_slots i32 [B], counts i32 [B,r_cap],
    slot_label i64 [B,r_cap], offsets i32 [B,r_cap+1], order i32 [B,P]]  (bit-exact with sppp.py:91-128)."""
    _cuda(labels)
    if labels.dtype != torch.int64 or labels.dim() != 3:
        raise ValueError("sppp_assign: labels must be an int64 [B,H,W] tensor")
    labels = labels.contiguous()
    B, Hh, Ww = labels.shape
    g = img_size // patch_size
    P = g * g
    dev = labels.device
    dom = torch.empty((B, P), dtype=torch.int64, device=dev)
    slot = torch.empty((B, P), dtype=torch.int32, device=dev)
    num_slots = torch.empty((B,), dtype=torch.int32, device=dev)
    counts = torch.empty((B, r_cap), dtype=torch.int32, device=dev)
    slot_label = torch.empty((B, r_cap), dtype=torch.int64, device=dev)
    offsets = torch.empty((B, r_cap + 1), dtype=torch.int32, device=dev)
    order = torch.empty((B, P), dtype=torch.int32, device=dev)
    if B * P:
        work = B * Hh * Ww * 8.0 + 2.0 * B * P * 4 + B * r_cap * 4
        rc = L.call("sppp_assign", work, L.lib().favit_sppp_assign, _p(labels), B, Hh, Ww, patch_size, g, _p(dom), _p(slot), _p(num_slots),
                                       _p(counts), _p(slot_label), _p(offsets), _p(order), r_cap, _stream())
        L.check(rc, "favit_sppp_assign")
    return [dom, slot, num_slots, counts, slot_label, offsets, order]


@sppp_assign.register_fake
def _(labels, patch_size, img_size, r_cap):
    B = labels.shape[0]
    P = (img_size // patch_size) ** 2
    i32, i64 = torch.int32, torch.int64
    return [labels.new_empty((B, P), dtype=i64), labels.new_empty((B, P), dtype=i32),
            labels.new_empty((B,), dtype=i32), labels.new_empty((B, r_cap), dtype=i32),
            labels.new_empty((B, r_cap), dtype=i64), labels.new_empty((B, r_cap + 1), dtype=i32),
            labels.new_empty((B, P), dtype=i32)]


@torch.library.custom_op("favit::sppp_assign_centroids", mutates_args=())
def sppp_assign_centroids(labels: Tensor, patch_size: int, img_size: int, r_cap: int, K: int) -> List[Tensor]:
    """sppp_assign + sppp_centroids with a single pass over the label map: the seven tensors of sppp_assign followed by
    the fp32 [B,K,2] centroids."""
    _cuda(labels)
    if labels.dtype != torch.int64 or labels.dim() != 3:
        raise ValueError("sppp_assign_centroids: labels must be an int64 [B,H,W] tensor")
    labels = labels.contiguous()
    B, Hh, Ww = labels.shape
    g = img_size // patch_size
    P = g * g
    dev = labels.device
    dom = torch.empty((B, P), dtype=torch.int64, device=dev)
    slot = torch.empty((B, P), dtype=torch.int32, device=dev)
    num_slots = torch.empty((B,), dtype=torch.int32, device=dev)
    counts = torch.empty((B, r_cap), dtype=torch.int32, device=dev)
    slot_label = torch.empty((B, r_cap), dtype=torch.int64, device=dev)
    offsets = torch.empty((B, r_cap + 1), dtype=torch.int32, device=dev)
    order = torch.empty((B, P), dtype=torch.int32, device=dev)
    cent = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
    if B * P and K:
        acc = torch.empty((B, 3, K), dtype=torch.int64, device=dev)
        work = B * Hh * Ww * 8.0 + 2.0 * B * P * 4 + B * r_cap * 4 + B * K * 8.0
        rc = L.call("sppp_assign", work, L.lib().favit_sppp_assign_centroids, _p(labels), B, Hh, Ww, patch_size, g, _p(dom),
                    _p(slot), _p(num_slots), _p(counts), _p(slot_label), _p(offsets), _p(order), r_cap, K, _p(acc),
                    _p(cent), _stream())
        L.check(rc, "favit_sppp_assign_centroids")
    return [dom, slot, num_slots, counts, slot_label, offsets, order, cent]


@sppp_assign_centroids.register_fake
def _(labels, patch_size, img_size, r_cap, K):
    B = labels.shape[0]
    P = (img_size // patch_size) ** 2
    i32, i64 = torch.int32, torch.int64
    return [labels.new_empty((B, P), dtype=i64), labels.new_empty((B, P), dtype=i32),
            labels.new_empty((B,), dtype=i32), labels.new_empty((B, r_cap), dtype=i32),
            labels.new_empty((B, r_cap), dtype=i64), labels.new_empty((B, r_cap + 1), dtype=i32),
            labels.new_empty((B, P), dtype=i32), labels.new_empty((B, K, 2), dtype=torch.float32)]


@torch.library.custom_op("favit::sppp_centroids", mutates_args=())
def sppp_centroids(labels: Tensor, K: int) -> Tensor:
    """labels int64 [B,H,W] -> fp32 [B,K,2] (x, y) centroids of labels 0..K-1 in normalised coordinates, (0.5, 0.5) for
    labels without pixels (sppp_mhla.py:226-262)."""
    _cuda(labels)
    if labels.dtype != torch.int64 or labels.dim() != 3:
        raise ValueError("sppp_centroids: labels must be an int64 [B,H,W] tensor")
    labels = labels.contiguous()
    B, Hh, Ww = labels.shape
    out = torch.empty((B, K, 2), dtype=torch.float32, device=labels.device)
    if B * K:
        acc = torch.empty((B, 3, K), dtype=torch.int64, device=labels.device)
        rc = L.call("sppp_centroids", B * Hh * Ww * 8.0 + B * K * 8.0, L.lib().favit_sppp_centroids, _p(labels), B, Hh, Ww, K,
                    _p(acc), _p(out), _stream())
        L.check(rc, "favit_sppp_centroids")
    return out


@sppp_centroids.register_fake
def _(labels, K):
    return labels.new_empty((labels.shape[0], K, 2), dtype=torch.float32)


@torch.library.custom_op("favit::sppp_pool_fwd", mutates_args=())
def sppp_pool_fwd(x: Tensor, order: Tensor, offsets: Tensor, num_slots: Tensor, R: int,
                  out_dtype: torch.dtype) -> Tensor:
    _cuda(x, order, offsets, num_slots)
    x = x.contiguous()
    B, P, D = x.shape
    r_cap = offsets.shape[1] - 1
    out = torch.empty((B, R, D), dtype=out_dtype, device=x.device)
    if out.numel():
        work = float(x.numel() * x.element_size() + B * P * 4 + out.numel() * out.element_size() + B * R * 4)
        rc = L.call("sppp_pool_fwd", work, L.lib().favit_sppp_pool_fwd, _p(x), _dt(x), _p(order), _p(offsets), _p(num_slots), _p(out), _DT[out_dtype],
                                         B, P, R, D, r_cap, _stream())
        L.check(rc, "favit_sppp_pool_fwd")
    return out


@sppp_pool_fwd.register_fake
def _(x, order, offsets, num_slots, R, out_dtype):
    return x.new_empty((x.shape[0], R, x.shape[2]), dtype=out_dtype)


@torch.library.custom_op("favit::sppp_pool_pixels", mutates_args=())
def sppp_pool_pixels(image: Tensor, order: Tensor, offsets: Tensor, num_slots: Tensor, patch_size: int, R: int,
                     out_dtype: torch.dtype) -> Tensor:
    """image fp32 [B,C,H,W] -> [B, R, patch*patch*C]: the mean RAW pixel patch of every superpixel slot, in the feature
    order of PatchEmbedding's Linear ('p1 p2 c').  Projecting it gives the pooled patch embeddings (sppp_mhla.py:281-300)
    with R instead of P rows per image.  No gradient flows to the image."""
    _cuda(image, order, offsets, num_slots)
    if image.dtype != torch.float32 or image.dim() != 4:
        raise ValueError("sppp_pool_pixels: image must be an fp32 [B,C,H,W] tensor")
    image = image.contiguous()
    B, Cc, Hh, Ww = image.shape
    g = Hh // patch_size
    r_cap = offsets.shape[1] - 1
    F = patch_size * patch_size * Cc
    out = torch.empty((B, R, F), dtype=out_dtype, device=image.device)
    if out.numel():
        work = float(B * Cc * (g * patch_size) ** 2 * 4 + out.numel() * out.element_size() + B * g * g * 4)
        rc = L.call("sppp_pool_pixels", work, L.lib().favit_sppp_pool_pixels, _p(image), B, Cc, Hh, Ww, patch_size, g,
                    _p(order), _p(offsets), _p(num_slots), _p(out), _DT[out_dtype], R, r_cap, _stream())
        L.check(rc, "favit_sppp_pool_pixels")
    return out


@sppp_pool_pixels.register_fake
def _(image, order, offsets, num_slots, patch_size, R, out_dtype):
    B, Cc = image.shape[0], image.shape[1]
    return image.new_empty((B, R, patch_size * patch_size * Cc), dtype=out_dtype)


@torch.library.custom_op("favit::patchify", mutates_args=())
def patchify(image: Tensor, patch_size: int, out_dtype: torch.dtype) -> Tensor:
    """image fp32 [B,C,S,S] -> [B, (S/p)^2, p*p*C] in out_dtype: the rearrangement of models/vit.py:38-39 + the cast to
    the GEMM operand type in one pass.  No gradient flows to the image."""
    _cuda(image)
    if image.dtype != torch.float32 or image.dim() != 4 or image.shape[2] != image.shape[3]:
        raise ValueError("patchify: image must be a square fp32 [B,C,S,S] tensor")
    image = image.contiguous()
    B, Cc, S, _ = image.shape
    g = S // patch_size
    out = torch.empty((B, g * g, patch_size * patch_size * Cc), dtype=out_dtype, device=image.device)
    if out.numel():
        rc = L.call("patchify", float(image.numel() * 4 + out.numel() * out.element_size()), L.lib().favit_patchify,
                    _p(image), _p(out), _DT[out_dtype], B, Cc, S, patch_size, _stream())
        L.check(rc, "favit_patchify")
    return out


@patchify.register_fake
def _(image, patch_size, out_dtype):
    B, Cc, S, _ = image.shape
    g = S // patch_size
    return image.new_empty((B, g * g, patch_size * patch_size * Cc), dtype=out_dtype)


@torch.library.custom_op("favit::sppp_embed_tokens", mutates_args=())
def sppp_embed_tokens(pooled: Tensor, cls_token: Tensor, centroids: Tensor) -> Tensor:
    """cat(cls, pooled) + dynamic positional encoding of the centroids (sppp_mhla.py:302-310, sppp.py:271-299):
    pooled fp32 [B,R,D], cls_token [1,1,D] or [D], centroids fp32 [B,R,2] -> fp32 [B,R+1,D]."""
    _cuda(pooled, cls_token, centroids)
    pooled = pooled.float().contiguous()
    cls = cls_token.float().reshape(-1).contiguous()
    cen = centroids.float().contiguous()
    B, R, D = pooled.shape
    if cls.numel() != D or cen.shape != (B, R, 2):
        raise ValueError(f"sppp_embed_tokens: pooled {tuple(pooled.shape)}, cls {tuple(cls_token.shape)}, centroids {tuple(cen.shape)}")
    out = torch.empty((B, R + 1, D), dtype=torch.float32, device=pooled.device)
    if out.numel():
        rc = L.call("sppp_embed", float(out.numel() * 8), L.lib().favit_sppp_embed_tokens, _p(pooled), _p(cls), _p(cen),
                    _p(out), B, R, D, _stream())
        L.check(rc, "favit_sppp_embed_tokens")
    return out


@sppp_embed_tokens.register_fake
def _(pooled, cls_token, centroids):
    B, R, D = pooled.shape
    return pooled.new_empty((B, R + 1, D), dtype=torch.float32)


def _embed_setup(ctx, inputs, output):
    ctx.cls_shape = inputs[1].shape


def _embed_backward(ctx, g):
    return g[:, 1:, :], g[:, 0, :].sum(dim=0).reshape(ctx.cls_shape), None


sppp_embed_tokens.register_autograd(_embed_backward, setup_context=_embed_setup)


@torch.library.custom_op("favit::sppp_pool_bwd", mutates_args=())
def sppp_pool_bwd(dout: Tensor, slot: Tensor, counts: Tensor, dx_dtype: torch.dtype) -> Tensor:
    _cuda(dout, slot, counts)
    dout = dout.contiguous()
    B, R, D = dout.shape
    P = slot.shape[1]
    r_cap = counts.shape[1]
    dx = torch.empty((B, P, D), dtype=dx_dtype, device=dout.device)
    if dx.numel():
        work = float(dout.numel() * dout.element_size() + B * P * 4 + dx.numel() * dx.element_size())
        rc = L.call("sppp_pool_bwd", work, L.lib().favit_sppp_pool_bwd, _p(dout), _dt(dout), _p(slot), _p(counts), _p(dx), _DT[dx_dtype], B, P, R, D,
                                         r_cap, _stream())
        L.check(rc, "favit_sppp_pool_bwd")
    return dx


@sppp_pool_bwd.register_fake
def _(dout, slot, counts, dx_dtype):
    return dout.new_empty((dout.shape[0], slot.shape[1], dout.shape[2]), dtype=dx_dtype)


@torch.library.custom_op("favit::sppp_pool", mutates_args=())
def sppp_pool(x: Tensor, slot: Tensor, counts: Tensor, order: Tensor, offsets: Tensor, num_slots: Tensor,
              R: int) -> Tensor:
    """Differentiable batched segment mean: x [B,P,D] -> fp32 [B,R,D] (the reference's output is always fp32,
    sppp.py:198)."""
    return sppp_pool_fwd(x, order, offsets, num_slots, R, torch.float32)


@sppp_pool.register_fake
def _(x, slot, counts, order, offsets, num_slots, R):
    return x.new_empty((x.shape[0], R, x.shape[2]), dtype=torch.float32)


def _pool_setup(ctx, inputs, output):
    x, slot, counts = inputs[0], inputs[1], inputs[2]
    ctx.save_for_backward(slot, counts)
    ctx.x_dtype = x.dtype


def _pool_backward(ctx, dout):
    slot, counts = ctx.saved_tensors
    return sppp_pool_bwd(dout, slot, counts, ctx.x_dtype), None, None, None, None, None, None


sppp_pool.register_autograd(_pool_backward, setup_context=_pool_setup)


# ---- 'max' and 'attention' pooling (models/sppp.py:178-184, 211-216) -----------------------------------------------
@torch.library.custom_op("favit::sppp_pool_max_fwd", mutates_args=())
def sppp_pool_max_fwd(x: Tensor, order: Tensor, offsets: Tensor, num_slots: Tensor, R: int) -> Tuple[Tensor, Tensor]:
    """x [B,P,D] -> (out fp32 [B,R,D], argmax int32 [B,R,D])."""
    _cuda(x, order, offsets, num_slots)
    x = x.contiguous()
    B, P, D = x.shape
    r_cap = offsets.shape[1] - 1
    out = torch.empty((B, R, D), dtype=torch.float32, device=x.device)
    arg = torch.empty((B, R, D), dtype=torch.int32, device=x.device)
    if out.numel():
        rc = L.call("sppp_pool_max", 0.0, L.lib().favit_sppp_pool_max_fwd, _p(x), _dt(x), _p(order), _p(offsets), _p(num_slots),
                    _p(out), _p(arg), B, P, R, D, r_cap, _stream())
        L.check(rc, "favit_sppp_pool_max_fwd")
    return out, arg


@sppp_pool_max_fwd.register_fake
def _(x, order, offsets, num_slots, R):
    B, P, D = x.shape
    return x.new_empty((B, R, D), dtype=torch.float32), x.new_empty((B, R, D), dtype=torch.int32)


@torch.library.custom_op("favit::sppp_pool_max_bwd", mutates_args=())
def sppp_pool_max_bwd(dout: Tensor, argmax: Tensor, P: int, dx_dtype: torch.dtype) -> Tensor:
    _cuda(dout, argmax)
    dout = dout.contiguous().float()
    B, R, D = dout.shape
    dx = torch.empty((B, P, D), dtype=dx_dtype, device=dout.device)
    if dx.numel():
        rc = L.call("sppp_pool_max", 0.0, L.lib().favit_sppp_pool_max_bwd, _p(dout), _p(argmax), _p(dx), _DT[dx_dtype], B, P, R, D,
                    _stream())
        L.check(rc, "favit_sppp_pool_max_bwd")
    return dx


@sppp_pool_max_bwd.register_fake
def _(dout, argmax, P, dx_dtype):
    return dout.new_empty((dout.shape[0], P, dout.shape[2]), dtype=dx_dtype)


@torch.library.custom_op("favit::sppp_pool_max", mutates_args=())
def sppp_pool_max(x: Tensor, order: Tensor, offsets: Tensor, num_slots: Tensor, R: int) -> Tuple[Tensor, Tensor]:
    """Differentiable batched segment max: x [B,P,D] -> (fp32 [B,R,D], argmax)."""
    return sppp_pool_max_fwd(x, order, offsets, num_slots, R)


@sppp_pool_max.register_fake
def _(x, order, offsets, num_slots, R):
    B, P, D = x.shape
    return x.new_empty((B, R, D), dtype=torch.float32), x.new_empty((B, R, D), dtype=torch.int32)


def _pool_max_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])
    ctx.P, ctx.x_dtype = inputs[0].shape[1], inputs[0].dtype
    ctx.set_materialize_grads(False)


def _pool_max_backward(ctx, dout, darg):
    (arg,) = ctx.saved_tensors
    if dout is None:
        return None, None, None, None, None
    return sppp_pool_max_bwd(dout, arg, ctx.P, ctx.x_dtype), None, None, None, None


sppp_pool_max.register_autograd(_pool_max_backward, setup_context=_pool_max_setup)


@torch.library.custom_op("favit::sppp_pool_attn_fwd", mutates_args=())
def sppp_pool_attn_fwd(x: Tensor, order: Tensor, offsets: Tensor, num_slots: Tensor, R: int) -> Tuple[Tensor, Tensor]:
    """x [B,P,D] -> (out fp32 [B,R,D], weights fp32 [B,P])."""
    _cuda(x, order, offsets, num_slots)
    x = x.contiguous()
    B, P, D = x.shape
    r_cap = offsets.shape[1] - 1
    out = torch.empty((B, R, D), dtype=torch.float32, device=x.device)
    w = torch.empty((B, P), dtype=torch.float32, device=x.device)
    if out.numel():
        rc = L.call("sppp_pool_attn", 0.0, L.lib().favit_sppp_pool_attn_fwd, _p(x), _dt(x), _p(order), _p(offsets), _p(num_slots),
                    _p(out), _p(w), B, P, R, D, r_cap, _stream())
        L.check(rc, "favit_sppp_pool_attn_fwd")
    return out, w


@sppp_pool_attn_fwd.register_fake
def _(x, order, offsets, num_slots, R):
    B, P, D = x.shape
    return x.new_empty((B, R, D), dtype=torch.float32), x.new_empty((B, P), dtype=torch.float32)


@torch.library.custom_op("favit::sppp_pool_attn_bwd", mutates_args=())
def sppp_pool_attn_bwd(x: Tensor, dout: Tensor, out: Tensor, weights: Tensor, order: Tensor, offsets: Tensor,
                       num_slots: Tensor) -> Tensor:
    _cuda(x, dout, out, weights, order, offsets, num_slots)
    x = x.contiguous()
    dout = dout.contiguous().float()
    B, P, D = x.shape
    R = out.shape[1]
    r_cap = offsets.shape[1] - 1
    dx = torch.empty_like(x)
    if dx.numel():
        rc = L.call("sppp_pool_attn", 0.0, L.lib().favit_sppp_pool_attn_bwd, _p(x), _dt(x), _p(dout), _p(out), _p(weights),
                    _p(order), _p(offsets), _p(num_slots), _p(dx), B, P, R, D, r_cap, _stream())
        L.check(rc, "favit_sppp_pool_attn_bwd")
    return dx


@sppp_pool_attn_bwd.register_fake
def _(x, dout, out, weights, order, offsets, num_slots):
    return torch.empty_like(x)


@torch.library.custom_op("favit::sppp_pool_attn", mutates_args=())
def sppp_pool_attn(x: Tensor, order: Tensor, offsets: Tensor, num_slots: Tensor, R: int) -> Tuple[Tensor, Tensor]:
    """Differentiable batched attention pooling: x [B,P,D] -> (fp32 [B,R,D], softmax weights [B,P])."""
    return sppp_pool_attn_fwd(x, order, offsets, num_slots, R)


@sppp_pool_attn.register_fake
def _(x, order, offsets, num_slots, R):
    B, P, D = x.shape
    return x.new_empty((B, R, D), dtype=torch.float32), x.new_empty((B, P), dtype=torch.float32)


def _pool_attn_setup(ctx, inputs, output):
    x, order, offsets, num_slots, R = inputs
    ctx.save_for_backward(x, output[0], output[1], order, offsets, num_slots)
    ctx.set_materialize_grads(False)


def _pool_attn_backward(ctx, dout, dw):
    x, out, w, order, offsets, num_slots = ctx.saved_tensors
    if dw is not None:
        raise RuntimeError("sppp_pool_attn: gradients through the returned weights are not supported")
    if dout is None:
        return None, None, None, None, None
    return sppp_pool_attn_bwd(x, dout, out, w, order, offsets, num_slots), None, None, None, None


sppp_pool_attn.register_autograd(_pool_attn_backward, setup_context=_pool_attn_setup)


# ------------------------------------------------------------------------------------------------
# GPU SLIC (models/sppp.py:26-74)
# ------------------------------------------------------------------------------------------------
def slic_grid(H: int, W: int, n_segments: int) -> Tuple[int, int]:
    import ctypes as C
    gy, gx = C.c_int(), C.c_int()
    L.check(L.lib().favit_slic_grid(H, W, n_segments, C.byref(gy), C.byref(gx)), "favit_slic_grid")
    return gy.value, gx.value


@torch.library.custom_op("favit::slic_segment", mutates_args=())
def slic_segment(image: Tensor, n_segments: int, compactness: float, sigma: float, iters: int) -> Tensor:
    """image fp32 [B,C,H,W] (C = 1 or 3) -> int64 [B,H,W] superpixel labels in [0, gy*gx)."""
    _cuda(image)
    if image.dtype != torch.float32 or image.dim() != 4:
        raise ValueError("slic_segment: image must be an fp32 [B,C,H,W] tensor")
    image = image.contiguous()
    B, Cc, Hh, Ww = image.shape
    gy, gx = slic_grid(Hh, Ww, n_segments)
    K = gy * gx
    dev = image.device
    labels = torch.empty((B, Hh, Ww), dtype=torch.int64, device=dev)
    if labels.numel():
        feat, tmp = torch.empty_like(image), torch.empty_like(image)
        centres = torch.empty((B, K, 2 + Cc), dtype=torch.float32, device=dev)
        sums = torch.empty((B, K, 3 + Cc), dtype=torch.int64, device=dev)
        work = float(image.numel() * 4) * (4 + 2 * (iters + 1)) + labels.numel() * 8.0 * (iters + 1)
        rc = L.call("slic", work, L.lib().favit_slic_segment, _p(image), B, Cc, Hh, Ww, n_segments, float(compactness),
                    float(sigma), int(iters), _p(labels), _p(feat), _p(tmp), _p(centres), _p(sums), _stream())
        L.check(rc, "favit_slic_segment")
    return labels


@slic_segment.register_fake
def _(image, n_segments, compactness, sigma, iters):
    return image.new_empty((image.shape[0], image.shape[2], image.shape[3]), dtype=torch.int64)
