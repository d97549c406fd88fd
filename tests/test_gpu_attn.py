"""GPU parity: favit::mhla_attn (window attention core, fwd + bwd) against the CPU oracle's closed form, which is
itself pinned to the reference by tests/test_oracle_golden.py.  Tolerances: 1e-4 relative fp32, 2e-2 bf16."""
import pytest
import torch

import oracle
from util import assert_close

pytestmark = pytest.mark.gpu

CASES = [
    # B, H, N, hd, W, mask
    (2, 3, 65, 64, 7, False),     # C1 shape
    (1, 2, 10, 64, 7, False),
    (2, 2, 5, 64, 7, False),      # N < W: every row pads with N-1 / 0
    (1, 1, 3, 64, 7, False),
    (2, 2, 1, 64, 7, False),      # N = 1
    (1, 3, 12, 32, 3, False),
    (1, 1, 6, 16, 1, False),      # W = 1
    (2, 2, 40, 64, 15, False),
    (1, 2, 17, 64, 31, False),    # C2 tokens with a window wider than the sequence
    (1, 12, 197, 64, 7, False),   # C4 shape, one image
    (2, 2, 10, 64, 7, True),
    (1, 2, 257, 64, 7, False),    # three chunks in the chunk-staged backward (N >= 64)
    (1, 2, 130, 64, 15, False),   # two chunks, wide window: halo + edge rows + 48 query slots per key tile
    (1, 1, 100, 32, 7, False),
    (1, 1, 72, 128, 3, False),
    (2, 1, 64, 64, 1, False),
    (1, 1, 128, 64, 7, False),    # chunk boundary == sequence end
    (1, 1, 129, 64, 7, False),    # a one-row second chunk
    (1, 2, 3, 64, 4, False),      # even window allowed when N <= W
    (1, 2, 4, 128, 4, False),
    # whole-sequence TMA kernels (head_dim 64, N <= 400): box / tile / CTA-packing edge cases
    (3, 5, 17, 64, 7, False),     # C2 tokens; 15 sequences, two per CTA -> the last CTA is half empty
    (1, 1, 16, 64, 7, False),     # exactly one tile, no padding rows
    (1, 2, 33, 64, 15, False),    # a one-row third tile under a wide window
    (2, 2, 256, 64, 7, False),    # the largest single TMA box
    (1, 2, 300, 64, 7, False),    # two boxes per operand, 19 warps
    (1, 1, 400, 64, 15, False),   # the largest supported sequence, 48 query slots per key tile
    (1, 2, 401, 64, 7, False),    # one past it: the chunked kernels (bf16) / per-warp staging (fp32)
]


def _run(B, H, N, hd, W, use_mask, dtype):
    from favit_b200 import ops
    torch.manual_seed(B * 1000 + N * 10 + W)
    qkv = torch.randn(B, N, 3, H, hd)
    dout = torch.randn(B, N, H * hd)
    mask = None
    if use_mask:
        mask = (torch.rand(B, N, N) > 0.3)
        idx = torch.arange(N)
        mask[:, idx, idx] = True
    qkv_d = qkv.to(dtype)           # quantise once so that both sides see the same inputs
    dout_d = dout.to(dtype)
    # oracle in fp64 on the CPU
    q64 = qkv_d.double().requires_grad_(True)
    q, k, v = [q64[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    o_ref, lse_ref = oracle.mhla_attn_core_closed_form(q, k, v, W, mask.double() if use_mask else None)
    o_ref = o_ref.permute(0, 2, 1, 3).reshape(B, N, H * hd)
    (o_ref * dout_d.double()).sum().backward()
    # CUDA path
    qc = qkv_d.cuda().requires_grad_(True)
    mc = mask.to(torch.uint8).cuda() if use_mask else None
    out, lse = ops.mhla_attn(qc, W, mc)
    out.backward(dout_d.cuda())
    assert out.dtype == dtype and lse.dtype == torch.float32
    assert_close(out, o_ref, dtype, "out")
    assert_close(lse, lse_ref, torch.float32 if dtype == torch.float32 else dtype, "lse")
    assert_close(qc.grad, q64.grad, dtype, "dqkv")
    scale = q64.grad.abs().max().item()      # dq is analytically 0 when every slot of a window is the same key (N = 1)
    for i, nm in enumerate("qkv"):
        assert_close(qc.grad[:, :, i], q64.grad[:, :, i], dtype, "d" + nm, factor=2.0, floor=1e-2 * scale)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B{}H{}N{}hd{}W{}{}".format(*c[:5], "m" if c[5] else ""))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_attn_core_matches_oracle(case, dtype):
    _run(*case, dtype)


def test_even_window_is_rejected_like_the_reference():
    from favit_b200 import ops
    qkv = torch.randn(1, 6, 3, 1, 64, device="cuda")
    with pytest.raises(RuntimeError):
        ops.mhla_attn(qkv, 4, None)


def test_fully_masked_row_is_nan_like_softmax():
    from favit_b200 import ops
    qkv = torch.randn(1, 8, 3, 1, 64, device="cuda")
    mask = torch.ones(1, 8, 8, dtype=torch.uint8, device="cuda")
    mask[0, 3, :] = 0
    out, _ = ops.mhla_attn(qkv, 3, mask)
    assert torch.isnan(out[0, 3]).all()
    assert torch.isfinite(out[0, :3]).all() and torch.isfinite(out[0, 4:]).all()


def test_large_shape_properties():
    """Full-size C4 slab: softmax rows are convex combinations -> |out| <= max |v| per head-dim; and the result does
    not depend on how the batch is split (images are independent)."""
    from favit_b200 import ops
    torch.manual_seed(1)
    B, N, H, hd, W = 64, 197, 12, 64, 7
    qkv = torch.randn(B, N, 3, H, hd, device="cuda", dtype=torch.bfloat16)
    out, lse = ops.mhla_attn(qkv, W, None)
    vmax = qkv[:, :, 2].float().abs().amax(dim=1, keepdim=True).reshape(B, 1, H * hd)
    assert (out.float().abs() <= vmax * 1.01 + 1e-3).all()
    out2, _ = ops.mhla_attn(qkv[5:9].contiguous(), W, None)
    assert torch.equal(out2, out[5:9])


@pytest.mark.parametrize("case", [(2, 3, 65, 64, 7), (1, 2, 5, 64, 7), (2, 2, 17, 64, 7), (1, 2, 40, 32, 15), (1, 1, 1, 64, 7),
                                  (1, 2, 197, 64, 7)], ids=lambda c: "B{}H{}N{}hd{}W{}".format(*c))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_attention_dropout_matches_gather_oracle_with_the_same_mask(case, dtype):
    """mhla.py:147: dropout on the [B,H,N,W] softmax output.  The kernels' counter-based keep-mask is reproduced on the
    CPU (oracle.dropout_keep_mask) and fed to the reference's gather formulation: forward and gradients must agree to
    the usual tolerance, duplicated edge slots included (each copy has its own Bernoulli)."""
    from favit_b200 import ops
    B, H, N, hd, W = case
    p, seed = 0.3, 123456789 + N
    torch.manual_seed(N * 7 + W)
    qkv_d = torch.randn(B, N, 3, H, hd).to(dtype)
    dout_d = torch.randn(B, N, H * hd).to(dtype)
    keep = torch.from_numpy(oracle.dropout_keep_mask(B, H, N, W, p, seed))
    assert 0.55 < keep.float().mean().item() < 0.85 or keep.numel() < 200
    q64 = qkv_d.double().requires_grad_(True)
    q, k, v = [q64[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    o_ref = oracle.mhla_attn_core_gather(q, k, v, W, None, keep, p).permute(0, 2, 1, 3).reshape(B, N, H * hd)
    (o_ref * dout_d.double()).sum().backward()
    qc = qkv_d.cuda().requires_grad_(True)
    out, _ = ops.mhla_attn(qc, W, None, p, seed)
    out.backward(dout_d.cuda())
    assert_close(out, o_ref, dtype, "out")
    assert_close(qc.grad, q64.grad, dtype, "dqkv", factor=2.0)
    # a different seed gives a different mask; p = 0 is the plain path
    out2, _ = ops.mhla_attn(qc.detach(), W, None, p, seed + 1)
    if N > 1:
        assert not torch.equal(out2, out.detach())


def test_module_attention_dropout_statistics():
    """MultiHeadLatentAttention(dropout=p).train(): E[output] over masks = the p = 0 output (inverted dropout), and the
    module runs forward + backward (reference: one p for attention-probability and projection dropout, mhla.py:43-44)."""
    from favit_b200.mhla import MultiHeadLatentAttention
    torch.manual_seed(0)
    m = MultiHeadLatentAttention(embed_dim=128, num_heads=2, window_size=7, dropout=0.25).cuda()
    x = torch.randn(2, 33, 128, device="cuda")
    m.eval()
    y0 = m(x)
    m.train()
    m.proj_dropout.p = 0.0                       # isolate the attention-probability dropout
    acc = torch.zeros_like(y0)
    n = 300
    for _ in range(n):
        acc += m(x).detach()
    err = (acc / n - y0).abs().max().item() / y0.abs().max().item()
    assert err < 0.08, err
    xg = x.clone().requires_grad_(True)
    m(xg).sum().backward()
    assert torch.isfinite(xg.grad).all() and m.latent_proj.weight.grad is not None


# wide windows: forward and backward on the tcgen05 / TMEM kernels (mhla_window_attn_tc.cu)
TC_CASES = [
    # B, H, N, W
    (2, 3, 197, 63),      # C4 tokens, two query tiles, the second one ragged
    (1, 2, 128, 63),      # exactly one tile
    (1, 2, 129, 31),      # one-row second tile
    (2, 2, 63, 63),       # N == W: every row but the middle one has duplicated-edge slots
    (1, 3, 64, 63),
    (3, 2, 65, 17),       # the narrowest window of this path
    (1, 2, 300, 33),      # two S chunks per warp at their limit (32 + 2h = 64 columns)
    (1, 2, 300, 35),      # three chunks
    (1, 1, 1025, 65),     # the widest window, nine tiles
    (1, 6, 4097, 63),     # C5-as-ViT tokens
]


@pytest.mark.parametrize("case", TC_CASES, ids=lambda c: "B{}H{}N{}W{}".format(*c))
def test_wide_window_tcgen05_forward_backward_match_oracle(case):
    from favit_b200 import _lib as L, ops
    B, H, N, W = case
    hd = 64
    torch.manual_seed(N + W)
    qkv = (torch.randn(B, N, 3, H, hd) * 1.5).to(torch.bfloat16)
    dout = torch.randn(B, N, H * hd).to(torch.bfloat16)
    qc = qkv.cuda().requires_grad_(True)
    out, lse = ops.mhla_attn_fwd(qc.detach(), W, None)
    assert L.last_kernel().startswith("attn_tc_fwd"), L.last_kernel()
    q64 = qkv.double().cuda().requires_grad_(True)          # the oracle's closed form, fp64, evaluated on the device
    q, k, v = [q64[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    o_ref, lse_ref = oracle.mhla_attn_core_closed_form(q, k, v, W, None)
    o_ref = o_ref.permute(0, 2, 1, 3).reshape(B, N, H * hd)
    assert_close(out, o_ref, torch.bfloat16, "out")
    assert_close(lse, lse_ref, torch.bfloat16, "lse")
    # every output row is a convex combination of V rows
    vmax = qkv[:, :, 2].float().abs().amax(dim=1).reshape(B, 1, H * hd).cuda()
    assert bool((out.float().abs() <= vmax * 1.01 + 1e-3).all())
    # backward on the tensor core as well: dQ query-major, dK / dV key-major, the duplicated-edge rows on the CUDA cores
    dqkv = ops.mhla_attn_bwd(qc.detach(), out, lse, dout.cuda(), W, None)
    assert L.last_kernel().startswith("attn_tc_bwd"), L.last_kernel()
    # the differentiable op (autograd runs the same kernels on its own thread)
    out2, _ = ops.mhla_attn(qc, W, None)
    out2.backward(dout.cuda())
    (o_ref * dout.double().cuda()).sum().backward()
    assert torch.equal(out2, out) and torch.equal(qc.grad, dqkv)
    scale = q64.grad.abs().max().item()
    for i, nm in enumerate("qkv"):
        assert_close(qc.grad[:, :, i], q64.grad[:, :, i], torch.bfloat16, "d" + nm, factor=2.0, floor=1e-2 * scale)


def test_wide_window_gather_oracle_agrees_on_a_small_case():
    """The closed form used above against the reference's own gather formulation (mhla.py:109-154) for a wide window."""
    torch.manual_seed(0)
    q, k, v = [torch.randn(1, 2, 70, 64, dtype=torch.float64) for _ in range(3)]
    a = oracle.mhla_attn_core_gather(q, k, v, 63)
    b, _ = oracle.mhla_attn_core_closed_form(q, k, v, 63)
    assert torch.allclose(a, b, rtol=1e-10, atol=1e-12)


# the chunked TMA kernels (mhla_window_attn_chunk.cu): forward for N > 48, backward for N > 400, windows <= 15
CHUNK_CASES = [
    # B, H, N, W
    (1, 2, 401, 7),       # the shortest sequence of this path: three chunks of nine tiles
    (2, 3, 1025, 7),      # 512-px ViT tokens: six chunks, the last one ragged
    (1, 12, 1025, 15),    # the widest window; 12 heads: column sums by global reductions
    (1, 2, 528, 15),      # the sequence ends exactly at a chunk end: the right halo tile is outside the sequence
    (1, 1, 4097, 3),      # 24 chunks
    (1, 2, 1024, 1),      # W = 1: no band beyond the diagonal, no duplicated edge keys
    (1, 1, 705, 7),       # a one-row last tile
    (2, 1, 417, 5),
    # 48 < N <= 400: the forward is chunked (64-row chunks: many small CTAs per SM), the backward is the whole-sequence
    # kernel
    (2, 3, 209, 7),
    (1, 2, 257, 15),
    (1, 12, 400, 7),
    (3, 3, 65, 7),        # C1 tokens: two chunks of three tiles, the second one holds one row
    (1, 12, 197, 7),      # C4 tokens: four chunks
    (1, 2, 49, 15),       # the shortest sequence of the chunked forward, the widest window: one chunk, both edge keys in it
    (1, 1, 64, 1),
]


@pytest.mark.parametrize("case", CHUNK_CASES, ids=lambda c: "B{}H{}N{}W{}".format(*c))
def test_long_sequence_chunked_forward_backward_match_oracle(case):
    from favit_b200 import _lib as L, ops, raw
    B, H, N, W = case
    hd = 64
    torch.manual_seed(N + W)
    qkv = (torch.randn(B, N, 3, H, hd) * 1.5).to(torch.bfloat16)
    dout = torch.randn(B, N, H * hd).to(torch.bfloat16)
    qc = qkv.cuda().requires_grad_(True)
    out, lse = ops.mhla_attn_fwd(qc.detach(), W, None)
    assert L.last_kernel().startswith("attn_chunk_fwd"), L.last_kernel()
    q64 = qkv.double().cuda().requires_grad_(True)          # the oracle's closed form, fp64, evaluated on the device
    q, k, v = [q64[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    o_ref, lse_ref = oracle.mhla_attn_core_closed_form(q, k, v, W, None)
    o_ref = o_ref.permute(0, 2, 1, 3).reshape(B, N, H * hd)
    assert_close(out, o_ref, torch.bfloat16, "out")
    assert_close(lse, lse_ref, torch.bfloat16, "lse")
    dqkv = ops.mhla_attn_bwd(qc.detach(), out, lse, dout.cuda(), W, None)
    assert L.last_kernel().startswith("attn_chunk_bwd" if N > 400 else "attn_seq_bwd"), L.last_kernel()
    out2, _ = ops.mhla_attn(qc, W, None)
    out2.backward(dout.cuda())
    (o_ref * dout.double().cuda()).sum().backward()
    assert torch.equal(out2, out) and torch.equal(qc.grad, dqkv)      # deterministic: no atomics on the data path
    scale = q64.grad.abs().max().item()
    for i, nm in enumerate("qkv"):
        assert_close(qc.grad[:, :, i], q64.grad[:, :, i], torch.bfloat16, "d" + nm, factor=2.0, floor=1e-2 * scale)
    # the rows next to the sequence ends carry the duplicated-edge terms: check them on their own, tightly scaled
    h = W // 2
    if h:
        for rows in (slice(0, h + 1), slice(N - h - 1, N)):
            ref = q64.grad[:, rows]
            assert_close(qc.grad[:, rows], ref, torch.bfloat16, "edge rows", factor=2.0, floor=1e-2 * ref.abs().max().item())
    # the fused qkv-bias gradient (column sums of the stored bf16 gradient, fp32 atomics across chunks)
    dqkv2, sums = raw.attn_bwd(qc.detach().reshape(B * N, 3 * H * hd), out, lse, dout.cuda(), B, N, H, hd, W)
    assert torch.equal(dqkv2.reshape(dqkv.shape), dqkv)
    want = dqkv.float().sum(dim=(0, 1)).reshape(-1)
    assert torch.allclose(sums, want, rtol=2e-3, atol=2e-3 * float(want.abs().max()) + 1e-3)
