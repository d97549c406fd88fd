"""The whole transformer block around MHLA as two torch.library ops (`favit::block_fwd`, `favit::block_bwd`) made only
of favit kernels — the "next" row 1 of SURVEY.md §8f on top of the hot path.

Replaces /root/reference/models/vit_mhla.py:77-109 (TransformerBlock.forward with use_mhla=True: norm1 -> MHLA ->
residual -> norm2 -> MLP -> residual; MLP = models/vit.py:125-139) and its autograd, for the cases every model in the
reference uses: no attention mask, dropout inactive.

Dataflow (M = B*N tokens; cd = compute dtype, bf16 under autocast or fp32):
  forward   x (fp32 residual stream) -LN1-> xn (cd) -GEMM+bias-> qkv (cd) -window attention-> o (cd)
            -GEMM+bias+residual(x)-> x2 (fp32) -LN2-> xn2 (cd) -GEMM+bias+GELU-> h (cd; pre-activation saved)
            -GEMM+bias+residual(x2)-> x3 (fp32)
  backward  every dgrad / wgrad is the tcgen05 GEMM (MN-major operands, no transposes), GELU' is fused into the fc2
            dgrad epilogue, the residual-gradient add and the bf16 operand copy are fused into the LayerNorm backward.
The latent projection is folded into the qkv / proj weights by the caller (mhla.fold_latent, a few [hd x hd] torch
matmuls whose autograd yields the latent_proj gradients).
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor

from . import _lib as L
from . import raw, rng


@torch.library.custom_op("favit::block_fwd", mutates_args=())
def block_fwd(x: Tensor, ln1_w: Tensor, ln1_b: Tensor, wqkv: Tensor, bqkv: Tensor, wproj: Tensor, bproj: Tensor,
              ln2_w: Tensor, ln2_b: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, B: int, N: int, H: int,
              window: int, eps1: float, eps2: float, p_drop: float = 0.0, seed: Optional[Tensor] = None,
              layer: int = 0) -> List[Tensor]:
    """x [B*N, D] fp32; GEMM weights already in the compute dtype (bf16 or fp32), biases / LN parameters fp32.
    Returns [x3, xn, mu1, rs1, qkv, o, lse, x2, xn2, mu2, rs2, hpre, h].
    p_drop > 0 (bf16 only): the MLP's two dropouts (vit.py:131-138) fused into the fc1 / fc2 epilogues, masks drawn from
    the device seed `seed` (int64 [1]) at the offsets of (layer, site); `h` is then the masked activation."""
    M, D = x.shape
    cd = wqkv.dtype
    hd = D // H
    L.ROLE = "ln"
    xn, mu1, rs1 = raw.ln_fwd(x, ln1_w, ln1_b, cd, eps1)
    L.ROLE = "qkv"
    qkv, _ = raw.linear_fwd(xn, wqkv, bqkv, None, cd)
    L.ROLE = "attn"
    o, lse = raw.attn_fwd(qkv, B, N, H, hd, window)
    L.ROLE = "proj"
    # the attention branch joins the fp32 residual stream inside the LayerNorm kernel (coalesced row accesses) rather
    # than in the GEMM epilogue (one thread per row): same bytes, moved to where they stream at HBM speed
    a, _ = raw.linear_fwd(o, wproj, bproj, None, cd)
    L.ROLE = "ln"
    xn2, mu2, rs2, x2 = raw.ln_fwd(x, ln2_w, ln2_b, cd, eps2, delta=a)
    L.ROLE = "fc1"
    d1 = (p_drop, seed, raw.drop_offset(layer, 1)) if p_drop > 0 else None
    d2 = (p_drop, seed, raw.drop_offset(layer, 2)) if p_drop > 0 else None
    h, hpre = raw.linear_fwd(xn2, w1, b1, None, cd, gelu=True, save_preact=True, drop=d1)
    L.ROLE = "fc2"
    x3, _ = raw.linear_fwd(h, w2, b2, x2, torch.float32, drop=d2)
    L.ROLE = ""
    return [x3, xn, mu1, rs1, qkv, o, lse, x2, xn2, mu2, rs2, hpre, h]


@block_fwd.register_fake
def _(x, ln1_w, ln1_b, wqkv, bqkv, wproj, bproj, ln2_w, ln2_b, w1, b1, w2, b2, B, N, H, window, eps1, eps2, p_drop=0.0,
      seed=None, layer=0):
    M, D = x.shape
    cd = wqkv.dtype
    f32 = torch.float32
    e = lambda shape, dt: x.new_empty(shape, dtype=dt)
    Hd = w1.shape[0]
    return [e((M, D), f32), e((M, D), cd), e((M,), f32), e((M,), f32), e((M, 3 * D), cd), e((M, D), cd),
            e((B, H, N), f32), e((M, D), f32), e((M, D), cd), e((M,), f32), e((M,), f32), e((M, Hd), cd),
            e((M, Hd), cd)]


@torch.library.custom_op("favit::block_bwd", mutates_args=())
def block_bwd(g: Tensor, x: Tensor, ln1_w: Tensor, wqkv: Tensor, wproj: Tensor, ln2_w: Tensor, w1: Tensor, w2: Tensor,
              xn: Tensor, mu1: Tensor, rs1: Tensor, qkv: Tensor, o: Tensor, lse: Tensor, x2: Tensor, xn2: Tensor,
              mu2: Tensor, rs2: Tensor, hpre: Tensor, h: Tensor, B: int, N: int, H: int, window: int,
              g_c: Optional[Tensor] = None, gsum: Optional[Tensor] = None, p_drop: float = 0.0,
              seed: Optional[Tensor] = None, layer: int = 0) -> List[Tensor]:
    """Returns [dx, small, dwqkv, dwproj, dw1, dw2, db2, dx_c] (fp32 except dx_c); `small` packs every small gradient of
    the block in one buffer (custom-op outputs may not alias each other): see `unpack_block_grads`.  `dx_c` is dx in
    the compute dtype (empty in fp32 mode) and the last D entries of `small` are its column sums: the LayerNorm
    backward produces both on the way, and the caller may hand them to the block below as `g_c` / `gsum` (the operand
    copy and column sums of THAT block's incoming gradient g, i.e. its fc2 bias gradient) — explicitly, and only when
    it knows that g is the very tensor they were derived from (see FusedBlockPrefoldedFn.backward)."""
    M, D = x.shape
    cd = wqkv.dtype
    hd = D // H
    bf = cd == torch.bfloat16
    g = g.contiguous()
    # every small accumulator of this backward (bias-gradient column sums, LayerNorm dgamma / dbeta) lives in one zeroed
    # buffer: one fill launch per block instead of four
    Hd = w2.shape[1]
    z = torch.zeros((Hd + 10 * D,), dtype=torch.float32, device=x.device)
    z_db1, z_ln2, z_attn = z[:Hd], z[Hd:Hd + 3 * D].view(3, D), z[Hd + 3 * D:Hd + 6 * D]
    z_ln1, z_db2 = z[Hd + 6 * D:Hd + 9 * D].view(3, D), z[Hd + 9 * D:]
    d1 = d2 = None
    if p_drop > 0:
        # the gradient reaches fc2 through the dropout that followed it: mask and rescale it while casting it into a
        # GEMM operand; its column sums are fc2's bias gradient.  (What the block above handed over is unmasked.)
        d1, d2 = (p_drop, seed, raw.drop_offset(layer, 1)), (p_drop, seed, raw.drop_offset(layer, 2))
        L.ROLE = "fc2"
        g_c, gsum = raw.dropout_cast(g, cd, d2, colsum=z_db2), z_db2
    elif g_c is None:
        g_c, gsum = (g.to(cd) if bf else g), None
    L.ROLE = "fc2"
    dw2, db2 = raw.linear_wgrad(g_c, h, want_bias=gsum is None)
    if gsum is not None:
        db2 = gsum
    dhpre, db1 = raw.linear_dgrad(g_c, w2, hpre, cd, colsum=True, zeroed=z_db1, drop=d1)   # fc1's bias gradient (GELU' epilogue)
    L.ROLE = "fc1"
    dw1, _ = raw.linear_wgrad(dhpre, xn2, want_bias=False)
    dxn2 = raw.linear_dgrad(dhpre, w1, None, cd)
    L.ROLE = "ln"
    g2, g2_b, dg2, dbt2, dbp = raw.ln_bwd(dxn2, x2, mu2, rs2, ln2_w, g, bf, zeroed=z_ln2)   # dbp = column sums of g2
    g2_c = g2_b if bf else g2
    L.ROLE = "proj"
    dwp, _ = raw.linear_wgrad(g2_c, o, want_bias=False)
    do = raw.linear_dgrad(g2_c, wproj, None, cd)
    L.ROLE = "attn"
    dqkv, dbq = raw.attn_bwd(qkv, o, lse, do, B, N, H, hd, window, zeroed=z_attn)   # qkv bias gradient on the way
    L.ROLE = "qkv"
    dwq, _ = raw.linear_wgrad(dqkv, xn, want_bias=False)
    dxn = raw.linear_dgrad(dqkv, wqkv, None, cd)
    L.ROLE = "ln"
    g0, g0_b, dg1, dbt1, g0sum = raw.ln_bwd(dxn, x, mu1, rs1, ln1_w, g2, bf, zeroed=z_ln1)
    L.ROLE = ""
    # db2 is either the wgrad's own output or the column sums handed over by the block above (a slice of THAT block's
    # buffer): a private copy keeps this op's outputs free of aliases
    return [g0, z, dwq, dwp, dw1, dw2, db2.clone() if gsum is not None else db2,
            g0_b if bf else g0.new_empty((0,), dtype=cd)]


def unpack_block_grads(outs, D: int, Hd: int):
    """block_bwd outputs -> (dx, dln1_w, dln1_b, dwqkv, dbqkv, dwproj, dbproj, dln2_w, dln2_b, dw1, db1, dw2, db2).
    Layout of `small`: [db1 (Hd) | dln2_w, dln2_b, dbproj (3 x D) | dbqkv (3D) | dln1_w, dln1_b, column sums of dx (3 x D) |
    db2 accumulator of the dropout path (D)]."""
    g0, z, dwq, dwp, dw1, dw2, db2 = outs[:7]
    db1 = z[:Hd]
    ln2 = z[Hd:Hd + 3 * D].view(3, D)
    dbq = z[Hd + 3 * D:Hd + 6 * D]
    ln1 = z[Hd + 6 * D:Hd + 9 * D].view(3, D)
    return g0, ln1[0], ln1[1], dwq, dbq, dwp, ln2[2], ln2[0], ln2[1], dw1, db1, dw2, db2


@block_bwd.register_fake
def _(g, x, ln1_w, wqkv, wproj, ln2_w, w1, w2, xn, mu1, rs1, qkv, o, lse, x2, xn2, mu2, rs2, hpre, h, B, N, H, window,
      g_c=None, gsum=None, p_drop=0.0, seed=None, layer=0):
    f32 = torch.float32
    e = lambda t: t.new_empty(t.shape, dtype=f32)
    D = x.shape[1]
    v = lambda n: x.new_empty((n,), dtype=f32)
    cd = wqkv.dtype
    dxc = x.new_empty(x.shape if cd == torch.bfloat16 else (0,), dtype=cd)
    return [e(x), v(w2.shape[1] + 10 * D), e(wqkv), e(wproj), e(w1), e(w2), v(D), dxc]


@torch.library.custom_op("favit::latent_fold_fwd", mutates_args=())
def latent_fold_fwd(qkv_w: Tensor, qkv_b: Tensor, proj_w: Tensor, proj_b: Tensor, lat_w: Tensor, lat_b: Tensor, H: int,
                    cd: torch.dtype) -> List[Tensor]:
    """[wqkv' (cd), bqkv' (fp32), wproj' (cd), bproj' (fp32)] with latent_proj folded in (mhla.py:105-106)."""
    return list(raw.fold_fwd(qkv_w, qkv_b, proj_w, proj_b, lat_w, lat_b, H, cd))


@latent_fold_fwd.register_fake
def _(qkv_w, qkv_b, proj_w, proj_b, lat_w, lat_b, H, cd):
    f32 = torch.float32
    return [qkv_w.new_empty(qkv_w.shape, dtype=cd), qkv_b.new_empty(qkv_b.shape, dtype=f32),
            proj_w.new_empty(proj_w.shape, dtype=cd), proj_b.new_empty(proj_b.shape, dtype=f32)]


@torch.library.custom_op("favit::latent_fold_bwd", mutates_args=("dwqkv", "dbqkv", "dwproj"))
def latent_fold_bwd(qkv_w: Tensor, qkv_b: Tensor, proj_w: Tensor, lat_w: Tensor, lat_b: Tensor, dwqkv: Tensor,
                    dbqkv: Tensor, dwproj: Tensor, dbproj: Tensor, H: int) -> List[Tensor]:
    """dwqkv / dbqkv / dwproj: gradients of the folded weights, rewritten IN PLACE into the gradients of qkv.weight,
    qkv.bias, proj.weight.  Returns [dlatent_w, dlatent_b]."""
    return list(raw.fold_bwd(qkv_w, qkv_b, proj_w, lat_w, lat_b, dwqkv, dbqkv, dwproj, dbproj, H))


@latent_fold_bwd.register_fake
def _(qkv_w, qkv_b, proj_w, lat_w, lat_b, dwqkv, dbqkv, dwproj, dbproj, H):
    return [lat_w.new_empty(lat_w.shape, dtype=torch.float32), lat_b.new_empty(lat_b.shape, dtype=torch.float32)]


class FusedBlockFn(torch.autograd.Function):
    """x [B,N,D] fp32 + the block's fp32 master parameters -> [B,N,D] fp32."""

    @staticmethod
    def forward(ctx, x, ln1_w, ln1_b, qkv_w, qkv_b, proj_w, proj_b, lat_w, lat_b, ln2_w, ln2_b, w1, b1, w2, b2, H,
                window, eps1, eps2, cd, p_drop=0.0):
        B, N, D = x.shape
        x2d = x.reshape(B * N, D)
        if not x2d.is_contiguous():
            x2d = x2d.contiguous()
        c = (lambda t: t.detach().to(cd)) if cd != torch.float32 else (lambda t: t.detach().contiguous())
        f = lambda t: t.detach().float().contiguous()
        raw_p = [f(qkv_w), f(qkv_b), f(proj_w), f(proj_b), f(lat_w), f(lat_b)]
        wq_c, bq, wp_c, bp = latent_fold_fwd(*raw_p, H, cd)
        w1_c, w2_c = c(w1), c(w2)
        seed = rng.call_seed(x.device) if p_drop > 0 else None
        outs = block_fwd(x2d.detach(), f(ln1_w), f(ln1_b), wq_c, bq, wp_c, bp, f(ln2_w), f(ln2_b), w1_c, f(b1), w2_c,
                         f(b2), B, N, H, window, eps1, eps2, p_drop, seed, 0)
        x3 = outs[0]
        ctx.save_for_backward(x2d, ln1_w, ln2_w, wq_c, wp_c, w1_c, w2_c, qkv_w, qkv_b, proj_w, lat_w, lat_b, *outs[1:])
        ctx.dims = (B, N, H, window)
        ctx.drop = (p_drop, seed)
        return x3.view(B, N, D)

    @staticmethod
    def backward(ctx, g):
        x2d, ln1_w, ln2_w, wq_c, wp_c, w1_c, w2_c, qkv_w, qkv_b, proj_w, lat_w, lat_b, *saved = ctx.saved_tensors
        B, N, H, window = ctx.dims
        D = x2d.shape[1]
        g2d = g.reshape(B * N, D)
        if g2d.dtype != torch.float32:
            g2d = g2d.float()
        f = lambda t: t.detach().float().contiguous()
        p_drop, seed = ctx.drop
        (dx, dln1_w, dln1_b, dwq, dbq, dwp, dbp, dln2_w, dln2_b, dw1, db1, dw2, db2) = unpack_block_grads(block_bwd(
            g2d, x2d, f(ln1_w), wq_c, wp_c, f(ln2_w), w1_c, w2_c, *saved, B, N, H, window, None, None, p_drop, seed, 0),
            D, w2_c.shape[1])
        # gradients of the folded qkv / proj weights -> qkv.weight, qkv.bias, proj.weight (in place) + latent_proj
        dlw, dlb = latent_fold_bwd(f(qkv_w), f(qkv_b), f(proj_w), f(lat_w), f(lat_b), dwq, dbq, dwp, dbp, H)
        return (dx.view(B, N, D), dln1_w, dln1_b, dwq, dbq, dwp, dbp, dlw, dlb, dln2_w, dln2_b, dw1, db1, dw2, db2,
                None, None, None, None, None, None)


class BatchedFoldFn(torch.autograd.Function):
    """latent_proj folded into qkv / proj for ALL blocks of a model in one call (raw.fold_fwd_batched).

    Inputs: 6 fp32 master parameters per layer (qkv.weight, qkv.bias, proj.weight, proj.bias, latent_proj.weight,
    latent_proj.bias).  Outputs: a one-element fp32 `token` and, per layer, (wqkv', bqkv', wproj', bproj') marked
    non-differentiable.  The gradients of the folded weights do not travel through autograd (they are fp32 while the
    folded weights are bf16): every block's backward leaves them in `stash`, and because only the FIRST block consumes
    `token`, autograd runs this node's backward after the last block backward of the pass, where one batched call maps
    all of them back to the master parameters."""

    @staticmethod
    def forward(ctx, H, cd, stash, nl, *tensors):
        """tensors: 6 fold parameters per layer, then (fc1.weight, fc2.weight) per layer, whose compute-dtype copies are
        made here as well (one cast launch for all blocks); their gradients come from the blocks, not from this node."""
        params, mlp = tensors[:6 * nl], tensors[6 * nl:]
        f = lambda t: t.detach().float().contiguous()
        layers = [tuple(f(t) for t in params[6 * i:6 * i + 6]) for i in range(nl)]
        folded = raw.fold_fwd_batched(layers, H, cd)
        flat = [t for lay in folded for t in lay]
        if cd == torch.bfloat16:
            casts = raw.cast_bf16_batched(list(mlp))
        else:
            casts = [t.detach().contiguous() for t in mlp]
        ctx.save_for_backward(*params)
        ctx.H, ctx.stash, ctx.n_mlp = H, stash, len(mlp)
        ctx.mark_non_differentiable(*flat, *casts)
        return (torch.zeros(1, dtype=torch.float32, device=params[0].device), *flat, *casts)

    @staticmethod
    def backward(ctx, gtoken, *unused):
        params = ctx.saved_tensors
        nl = len(params) // 6
        f = lambda t: t.detach().float().contiguous()
        layers = [tuple(f(params[6 * i + k]) for k in (0, 1, 2, 4, 5)) for i in range(nl)]
        grads = [ctx.stash.pop(i) for i in range(nl)]
        dl = raw.fold_bwd_batched(layers, grads, ctx.H)
        out = []
        for (dwq, dbq, dwp, dbp), (dlw, dlb) in zip(grads, dl):
            out += [dwq, dbq, dwp, dbp, dlw, dlb]
        return (None, None, None, None, *out, *([None] * ctx.n_mlp))


class FusedBlockPrefoldedFn(torch.autograd.Function):
    """FusedBlockFn with the folded qkv / proj weights supplied by BatchedFoldFn (see there for `token` / `stash`)."""

    @staticmethod
    def forward(ctx, x, token, ln1_w, ln1_b, ln2_w, ln2_b, w1, b1, w2, b2, wq_c, bq, wp_c, bp, w1_c, w2_c, layer, stash, H,
                window, eps1, eps2, cd, p_drop=0.0, seed=None):
        B, N, D = x.shape
        x2d = x.reshape(B * N, D)
        if not x2d.is_contiguous():
            x2d = x2d.contiguous()
        f = lambda t: t.detach().float().contiguous()
        outs = block_fwd(x2d.detach(), f(ln1_w), f(ln1_b), wq_c, bq, wp_c, bp, f(ln2_w), f(ln2_b), w1_c, f(b1), w2_c,
                         f(b2), B, N, H, window, eps1, eps2, p_drop, seed, layer)
        ctx.save_for_backward(x2d, ln1_w, ln2_w, wq_c, wp_c, w1_c, w2_c, *outs[1:])
        ctx.dims = (B, N, H, window)
        ctx.drop = (p_drop, seed)
        ctx.layer, ctx.stash, ctx.has_token = layer, stash, token is not None
        return outs[0].view(B, N, D)

    @staticmethod
    def backward(ctx, g):
        x2d, ln1_w, ln2_w, wq_c, wp_c, w1_c, w2_c, *saved = ctx.saved_tensors
        B, N, H, window = ctx.dims
        D = x2d.shape[1]
        g2d = g.reshape(B * N, D)
        if g2d.dtype != torch.float32:
            g2d = g2d.float()
        f = lambda t: t.detach().float().contiguous()
        # Hand-over from the block above (layer + 1), whose backward ran just before this one in the chain run_blocks
        # built: the compute-dtype copy of its dx and the column sums of dx, both by-products of its LayerNorm backward.
        # They are used only if the gradient autograd delivers here IS that dx, unmodified: same storage (the entry
        # keeps dx alive, so the address cannot have been recycled) and same version counter (autograd accumulates
        # fan-out gradients in place, which bumps it).  Anything else (a module between the blocks, a second consumer
        # of the block's input) falls back to casting g.
        g_c = gsum = None
        handed = ctx.stash.pop(("handover", ctx.layer + 1), None)
        if handed is not None:
            dx_above, ver, dx_c, dxsum = handed
            if (g2d.data_ptr() == dx_above.data_ptr() and g2d.shape == dx_above.shape and g2d.is_contiguous()
                    and g2d._version == ver and dx_above._version == ver):
                g_c, gsum = (dx_c if dx_c.numel() else g2d), dxsum
        p_drop, seed = ctx.drop
        outs = block_bwd(g2d, x2d, f(ln1_w), wq_c, wp_c, f(ln2_w), w1_c, w2_c, *saved, B, N, H, window, g_c, gsum, p_drop,
                         seed, ctx.layer)
        (dx, dln1_w, dln1_b, dwq, dbq, dwp, dbp, dln2_w, dln2_b, dw1, db1, dw2, db2) = unpack_block_grads(
            outs, D, w2_c.shape[1])
        if ctx.layer > 0:
            ctx.stash[("handover", ctx.layer)] = (dx, dx._version, outs[7], outs[1][-2 * D:-D])
        ctx.stash[ctx.layer] = (dwq, dbq, dwp, dbp)
        gtok = torch.zeros(1, dtype=torch.float32, device=dx.device) if ctx.has_token else None
        return (dx.view(B, N, D), gtok, dln1_w, dln1_b, dln2_w, dln2_b, dw1, db1, dw2, db2, None, None, None, None, None,
                None, None, None, None, None, None, None, None, None, None)


def run_blocks(blocks, x: Tensor, compute_dtype: torch.dtype, p_drop: float = 0.0) -> Tensor:
    """x through a list of TransformerBlock-like modules (attributes norm1, attn, norm2, mlp.fc1, mlp.fc2), every one of
    which the caller found `fusable`: the latent fold of all blocks is one launch set per pass instead of one per block."""
    attn0 = blocks[0].attn
    H, window = attn0.num_heads, attn0.window_size
    params, mlp = [], []
    for blk in blocks:
        a = blk.attn
        params += [a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias, a.latent_proj.weight, a.latent_proj.bias]
        mlp += [blk.mlp.fc1.weight, blk.mlp.fc2.weight]
    stash, nl = {}, len(blocks)
    token, *flat = BatchedFoldFn.apply(H, compute_dtype, stash, nl, *params, *mlp)
    if not token.requires_grad:
        token = None
    casts = flat[4 * nl:]
    seed = rng.call_seed(x.device) if p_drop > 0 else None     # one seed per call; (layer, site) offsets tell masks apart
    for i, blk in enumerate(blocks):
        wq_c, bq, wp_c, bp = flat[4 * i:4 * i + 4]
        fc1, fc2 = blk.mlp.fc1, blk.mlp.fc2
        x = FusedBlockPrefoldedFn.apply(x, token if i == 0 else None, blk.norm1.weight, blk.norm1.bias, blk.norm2.weight,
                                        blk.norm2.bias, fc1.weight, fc1.bias, fc2.weight, fc2.bias, wq_c, bq, wp_c, bp,
                                        casts[2 * i], casts[2 * i + 1], i, stash, H, window, blk.norm1.eps, blk.norm2.eps,
                                        compute_dtype, p_drop, seed)
    return x


def fused_block(x, ln1, attn, ln2, fc1, fc2, compute_dtype: torch.dtype, p_drop: float = 0.0):
    """attn: a MultiHeadLatentAttention module (qkv / proj / latent_proj parameters are read directly).
    p_drop: the MLP dropout probability when it is active (training mode), else 0."""
    return FusedBlockFn.apply(x, ln1.weight, ln1.bias, attn.qkv.weight, attn.qkv.bias, attn.proj.weight,
                              attn.proj.bias, attn.latent_proj.weight, attn.latent_proj.bias, ln2.weight, ln2.bias,
                              fc1.weight, fc1.bias, fc2.weight, fc2.bias, attn.num_heads, attn.window_size, ln1.eps,
                              ln2.eps, compute_dtype, p_drop)


def fusable(x: Tensor, attn, mlp_dropout_p: float, training: bool, attention_mask, compute_dtype, hidden: int) -> bool:
    """Conditions under which the block runs as the fused ops (otherwise the caller composes the unfused ops)."""
    if attention_mask is not None or not x.is_cuda or x.dtype != torch.float32 or x.dim() != 3:
        return False
    if training and attn.attn_dropout.p > 0:      # attention-probability AND projection dropout (mhla.py:43-44): unfused
        return False
    if training and mlp_dropout_p > 0 and compute_dtype != torch.bfloat16:
        return False                              # the fused dropout epilogues exist on the bf16 path only
    D = x.shape[-1]
    if compute_dtype not in (torch.bfloat16, torch.float32):
        return False
    if attn.head_dim not in (16, 32, 64) or D % 8 or D > 1024 or hidden % 8:
        return False
    if attn.window_size % 2 == 0 and x.shape[1] > attn.window_size:
        return False
    return True
