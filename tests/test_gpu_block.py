"""GPU parity of the fused transformer block (favit::block_fwd / block_bwd: LayerNorm, GEMMs with fused
bias/GELU/residual, window attention) against the CPU oracle block, itself pinned to the reference by the golden
fixtures; plus the LayerNorm kernels against torch."""
import pytest
import torch

import oracle
from oracle.models_oracle import block as oracle_block
from util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,D", [(37, 64), (300, 192), (1000, 384), (520, 768), (64, 1024)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_layernorm_fwd_bwd(M, D, out_dtype):
    from favit_b200 import raw
    torch.manual_seed(D)
    x = (torch.randn(M, D, device="cuda") * 2 + 0.5)
    gamma = torch.randn(D, device="cuda")
    beta = torch.randn(D, device="cuda")
    y, mean, rstd = raw.ln_fwd(x, gamma, beta, out_dtype, 1e-5)
    xr = x.double().requires_grad_(True)
    gr, br = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-5)
    assert_close(y, yr, out_dtype, "ln y")
    assert_close(mean, xr.mean(dim=1), torch.float32, "mean")
    dy = torch.randn(M, D, device="cuda").to(out_dtype)
    dres = torch.randn(M, D, device="cuda")
    yr.backward(dy.double())
    dx, dxb, dg, db, dxs = raw.ln_bwd(dy, x, mean, rstd, gamma, dres, True)
    assert_close(dxs, (xr.grad + dres.double()).sum(dim=0), torch.float32, "dx column sums", factor=5.0,
                 floor=1e-3 * float(dx.abs().max()) * M ** 0.5)
    assert_close(dx, xr.grad + dres.double(), torch.float32, "ln dx", factor=3.0)
    assert_close(dxb, xr.grad + dres.double(), torch.bfloat16, "ln dx bf16")
    assert_close(dg, gr.grad, torch.float32, "dgamma", factor=5.0)
    assert_close(db, br.grad, torch.float32, "dbeta", factor=5.0)


def _block_case(mode, B, N, D, H, W, ratio, seq_block):
    from favit_b200.mhla import MHLATransformerBlock
    from favit_b200.models import TransformerBlock
    torch.manual_seed(B * N + D)
    if seq_block:
        blk = MHLATransformerBlock(embed_dim=D, num_heads=H, window_size=W, mlp_ratio=ratio)
    else:
        blk = TransformerBlock(embed_dim=D, num_heads=H, mlp_ratio=ratio, window_size=W, use_mhla=True)
    for p in blk.parameters():          # non-trivial LayerNorm / bias values
        if p.dim() == 1:
            torch.nn.init.normal_(p, mean=0.3, std=0.5)
    blk = blk.cuda()
    x = torch.randn(B, N, D)
    g = torch.randn(B, N, D)
    dtype = torch.float32 if mode == "fp32" else torch.bfloat16
    xc = x.cuda().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
        y = blk(xc)
    assert y.dtype == torch.float32          # fp32 residual stream, like the reference under autocast
    y.backward(g.cuda())
    sd = {k: v.detach().cpu().double().requires_grad_(True) for k, v in blk.state_dict().items()}
    xr = x.double().requires_grad_(True)
    yr = oracle_block(xr, sd, "", H, W)
    yr.backward(g.double())
    assert_close(y, yr, dtype, "block y")
    assert_close(xc.grad, xr.grad, dtype, "block dx", factor=2.0)
    for k, p in blk.named_parameters():
        assert_close(p.grad, sd[k].grad, dtype, f"block d{k}", factor=3.0)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("B,N,D,H,W,ratio,seq", [(2, 17, 128, 2, 7, 4.0, False), (3, 65, 192, 3, 7, 4.0, False),
                                                 (2, 10, 64, 1, 3, 2.0, True), (1, 197, 768, 12, 7, 4.0, False),
                                                 (2, 5, 128, 2, 7, 4.0, True)])
def test_fused_block_matches_oracle(mode, B, N, D, H, W, ratio, seq):
    from favit_b200 import fused_block
    calls = []
    orig = fused_block.fused_block
    fused_block.fused_block = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        _block_case(mode, B, N, D, H, W, ratio, seq)
    finally:
        fused_block.fused_block = orig
    assert calls, "the fused path was not taken"


def test_fused_and_unfused_paths_agree():
    """The unfused composition (mask / dropout cases) and the fused ops are the same math."""
    from favit_b200.models import TransformerBlock
    torch.manual_seed(5)
    blk = TransformerBlock(embed_dim=128, num_heads=2, window_size=7, use_mhla=True).cuda()
    x = torch.randn(2, 33, 128, device="cuda")
    y_fused = blk(x)
    mask = torch.ones(2, 33, 33, device="cuda")      # an all-ones mask forces the unfused path
    y_unfused = blk(x, mask)
    assert_close(y_unfused, y_fused, torch.float32, "fused vs unfused")


def test_batched_latent_fold_matches_per_layer_fold():
    """favit_latent_fold_{fwd,bwd}_batched (one launch set for all blocks) against the single-layer entry points."""
    from favit_b200 import raw
    torch.manual_seed(7)
    L, H, hd = 5, 3, 64
    D = H * hd
    dev = "cuda"
    layers = [(torch.randn(3 * D, D, device=dev) * 0.05, torch.randn(3 * D, device=dev) * 0.1,
               torch.randn(D, D, device=dev) * 0.05, torch.randn(D, device=dev) * 0.1,
               torch.eye(hd, device=dev) + torch.randn(hd, hd, device=dev) * 0.05, torch.randn(hd, device=dev) * 0.1)
              for _ in range(L)]
    for cd in (torch.bfloat16, torch.float32):
        batched = raw.fold_fwd_batched(layers, H, cd)
        for lay, got in zip(layers, batched):
            ref = raw.fold_fwd(*lay, H, cd)
            for g, r in zip(got, ref):   # the folded proj bias is summed over heads with fp32 atomics: last-bit order effects
                assert g.dtype == r.dtype and torch.allclose(g.float(), r.float(), rtol=1e-6, atol=1e-6)
    grads = [(torch.randn(3 * D, D, device=dev), torch.randn(3 * D, device=dev), torch.randn(D, D, device=dev),
              torch.randn(D, device=dev)) for _ in range(L)]
    ref_in = [tuple(t.clone() for t in g) for g in grads]
    bl = [(q, qb, p, lw, lb) for (q, qb, p, pb, lw, lb) in layers]
    dl = raw.fold_bwd_batched(bl, grads, H)
    for i in range(L):
        dlw, dlb = raw.fold_bwd(*bl[i], *ref_in[i], H)
        # dlat is accumulated with floating-point atomics: order-dependent in the last bits
        assert torch.allclose(dl[i][0], dlw, rtol=1e-5, atol=1e-4) and torch.allclose(dl[i][1], dlb, rtol=1e-5, atol=1e-4)
        for g, r in zip(grads[i][:3], ref_in[i][:3]):    # rewritten in place by both
            assert torch.equal(g, r)


def _stack_case(mid, towers):
    """Blocks run through models.run_blocks (batched fold + the explicit bf16-gradient hand-over between consecutive
    blocks) in arrangements where the gradient that reaches a block is NOT the tensor the block above produced."""
    from favit_b200.models import TransformerBlock, run_blocks
    torch.manual_seed(11)
    B, N, D, H, W = 2, 17, 128, 2, 7
    mk = lambda: torch.nn.ModuleList([TransformerBlock(embed_dim=D, num_heads=H, window_size=W, use_mhla=True)
                                      for _ in range(2)]).cuda()
    stacks = [mk() for _ in range(towers)]
    tail = mk()
    x = torch.randn(B, N, D)
    g = torch.randn(B, N, D)

    def forward(xin, blocks_fn, scale):
        outs = [blocks_fn(s, xin) for s in stacks]
        y = outs[0] if towers == 1 else outs[0] + outs[1]
        if mid:                                  # two element-wise ops between the fused stacks
            y = y * scale + 0.25
        return blocks_fn(tail, y)

    xc = x.cuda().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = forward(xc, lambda s, t: run_blocks(s, t), 1.5)
    y.backward(g.cuda())
    xr = x.double().requires_grad_(True)

    def oracle_stack(s, t):
        sd = {k: v.detach().cpu().double().requires_grad_(True) for k, v in s.state_dict().items()}
        oracle_stack.sds.append((s, sd))
        for i in range(len(s)):
            t = oracle_block(t, sd, f"{i}.", H, W)
        return t
    oracle_stack.sds = []
    yr = forward(xr, oracle_stack, 1.5)
    yr.backward(g.double())
    assert_close(y, yr, torch.bfloat16, "stack y")
    assert_close(xc.grad, xr.grad, torch.bfloat16, "stack dx", factor=2.0)
    for s, sd in oracle_stack.sds:
        for k, p in s.named_parameters():
            assert_close(p.grad, sd[k].grad, torch.bfloat16, f"stack d{k}", factor=3.0,
                         floor=1e-3 * float(xr.grad.abs().max()))


@pytest.mark.parametrize("mid,towers", [(False, 1), (True, 1), (False, 2), (True, 2)],
                         ids=["chain", "ops_between", "two_towers", "two_towers_ops_between"])
def test_block_gradient_handover_is_tied_to_the_tensor(mid, towers):
    _stack_case(mid, towers)


def test_second_backward_without_forward_reuses_nothing_stale():
    """backward(retain_graph=True) twice with different output gradients: the second pass must not pick up the
    operand copy / column sums the first pass left behind."""
    from favit_b200.models import TransformerBlock, run_blocks
    torch.manual_seed(12)
    B, N, D, H, W = 2, 17, 128, 2, 7
    blocks = torch.nn.ModuleList([TransformerBlock(embed_dim=D, num_heads=H, window_size=W, use_mhla=True)
                                  for _ in range(3)]).cuda()
    x = torch.randn(B, N, D, device="cuda", requires_grad=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = run_blocks(blocks, x)
    g1, g2 = torch.randn_like(y), torch.randn_like(y)
    y.backward(g1, retain_graph=True)
    first = {k: p.grad.clone() for k, p in blocks.named_parameters()}
    dx1 = x.grad.clone()
    for p in blocks.parameters():
        p.grad = None
    x.grad = None
    y.backward(g2, retain_graph=True)
    second = {k: p.grad.clone() for k, p in blocks.named_parameters()}
    dx2 = x.grad.clone()
    # linearity of backward in the output gradient: a third pass with g1 + g2 equals the sum of the two
    for p in blocks.parameters():
        p.grad = None
    x.grad = None
    y.backward(g1 + g2)
    assert_close(x.grad, dx1 + dx2, torch.bfloat16, "dx linearity")
    scale = float(max(v.abs().max() for v in first.values()))
    for k, p in blocks.named_parameters():
        assert_close(p.grad, first[k] + second[k], torch.bfloat16, f"linearity d{k}", floor=1e-2 * scale)
