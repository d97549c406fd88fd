"""GPU parity of the code paths the BENCHMARK runs, at the benchmark's own shapes (BASELINE configs C4 = ViT-B/16 with
50 432 tokens per GPU, C2 = SPPP ViT-S with 256 images, C5 = 512 px / patch 8 per-GPU share):

  * every epilogue of `gemm_bf16_tcgen05_2cta_kernel` (plain bf16, GELU dual store, GELU' + column sums, fp32 split-K
    reduce-add) through favit_linear_fwd / dgrad / wgrad, replacing nn.Linear at /root/reference/models/mhla.py:100,158
    and models/vit.py:125-139 — each test ASSERTS through favit_last_kernel() that the CTA-pair kernel is what ran, so
    that a change of the dispatch rule cannot silently move a test to another kernel;
  * the persistent multi-item ring of `sppp_pool_fwd_tma_kernel` (1 536 items on 592 CTAs at C2) and the CSR-stream
    kernel at C5, forward and backward (models/sppp.py:192-223, models/sppp_mhla.py:283-300);
  * one full-width training step (D 768, 12 heads, 197 tokens, 256 images) replayed from the CUDA graph that
    engine.TrainStep captures, against the oracle's loss and every parameter gradient.

The checker is torch fp64 (plain matmuls, and oracle/ for the model and the pooling) evaluated on the GPU: at these sizes the
fp64 reference needs tens of GFLOP..TFLOP, seconds on the device and minutes on the host cores.
"""
import numpy as np
import pytest
import torch

import oracle
from util import assert_close, rel_err

pytestmark = pytest.mark.gpu

M4 = 256 * 197            # C4 tokens per GPU
M2 = 256 * 17             # C2 tokens per GPU


def _ran_2cta(aux=None, **fields):
    from favit_b200 import _lib as L
    k = L.last_kernel()
    assert k.startswith("gemm_bf16_tcgen05_2cta_kernel"), f"expected the CTA-pair kernel, dispatch chose: {k!r}"
    if aux is not None:
        assert f"<AUX={int(aux)}>" in k, k
    for name, val in fields.items():
        assert f" {name}={val}" in k, (name, val, k)
    return k


def _bf(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


# the forward GEMMs of a ViT-B block at C4 and those of the ViT-S block at C2 that qualify for the CTA-pair kernel
FWD = [("c4_qkv", M4, 768, 2304), ("c4_proj", M4, 768, 768), ("c4_fc1", M4, 768, 3072), ("c4_fc2", M4, 3072, 768),
       ("c2_fc1", M2, 384, 1536)]


@pytest.mark.parametrize("name,M,K,N", FWD, ids=[f[0] for f in FWD])
def test_2cta_forward_bias_bf16(name, M, K, N):
    """Y = X.W^T + b, bf16 out (qkv / proj as block_fwd issues them)."""
    from favit_b200 import ops
    x, w = _bf(M, K, seed=1), _bf(N, K, scale=0.05, seed=2)
    b = torch.randn(N, device="cuda")
    y, _ = ops.linear_fwd(x, w, b, None, False, torch.bfloat16, False)
    _ran_2cta(aux=0, act=0, c_fp32=0, reduce=0)
    ref = x.double() @ w.double().t() + b.double()
    assert_close(y, ref, torch.bfloat16, name)
    # bf16 rounding of an fp32 accumulator: the error of every element is below one bf16 ulp of its own magnitude
    err = (y.double() - ref).abs()
    assert bool((err <= ref.abs() * 2 ** -7 + 1e-3).all()), name


@pytest.mark.parametrize("name,M,K,N", [FWD[2], FWD[4]], ids=["c4_fc1", "c2_fc1"])
def test_2cta_forward_gelu_dual_store(name, M, K, N):
    """fc1: pre-activation and GELU(pre-activation), two bf16 tensors from one accumulator tile (vit.py:125-139)."""
    from favit_b200 import ops
    x, w = _bf(M, K, seed=3), _bf(N, K, scale=0.06, seed=4)
    b = torch.randn(N, device="cuda") * 0.5
    y, pre = ops.linear_fwd(x, w, b, None, True, torch.bfloat16, True)
    _ran_2cta(aux=0, act=1, c_fp32=0)
    pre_ref = x.double() @ w.double().t() + b.double()
    assert_close(pre, pre_ref, torch.bfloat16, "preact")
    assert_close(y, torch.nn.functional.gelu(pre_ref), torch.bfloat16, "gelu")


def test_2cta_forward_fp32_out():
    """fp32 C through the 32-column staging units (plain store, no split)."""
    from favit_b200 import ops
    M, K, N = M4, 768, 768
    x, w = _bf(M, K, seed=5), _bf(N, K, scale=0.05, seed=6)
    b = torch.randn(N, device="cuda")
    y, _ = ops.linear_fwd(x, w, b, None, False, torch.float32, False)
    _ran_2cta(aux=0, act=0, c_fp32=1, reduce=0)
    assert rel_err(y, x.double() @ w.double().t() + b.double()) < 1e-5


def test_fc2_forward_with_residual_at_c4():
    """fc2 + bias + fp32 residual -> fp32 (the block's last GEMM) on the CTA-pair kernel; a ragged N (not a multiple of
    32) keeps the single-CTA kernel."""
    from favit_b200 import _lib as L, ops
    M, K, N = M4, 3072, 768
    x, w = _bf(M, K, seed=7), _bf(N, K, scale=0.03, seed=8)
    b = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    y, _ = ops.linear_fwd(x, w, b, res, False, torch.float32, False)
    _ran_2cta(aux=0, act=0, c_fp32=1, reduce=0, residual=1)
    assert rel_err(y, x.double() @ w.double().t() + b.double() + res.double()) < 1e-5
    N2 = 776
    w2, b2, res2 = _bf(N2, K, scale=0.03, seed=9), torch.randn(N2, device="cuda"), torch.randn(M, N2, device="cuda")
    y2, _ = ops.linear_fwd(x, w2, b2, res2, False, torch.float32, False)
    assert L.last_kernel().startswith("gemm_bf16_tcgen05_kernel<BN=256>"), L.last_kernel()
    assert rel_err(y2, x.double() @ w2.double().t() + b2.double() + res2.double()) < 1e-5


DGRAD = [("c4_fc1", M4, 3072, 768), ("c4_proj", M4, 768, 768), ("c4_qkv", M4, 2304, 768), ("c2_fc1", M2, 1536, 384)]


@pytest.mark.parametrize("name,M,N,K", DGRAD, ids=[d[0] for d in DGRAD])
def test_2cta_dgrad_plain_and_colsum(name, M, N, K):
    """dX = dY.W (MN-major B operand), bf16, with and without the fused column sums of the stored dX."""
    from favit_b200 import raw
    dy, w = _bf(M, N, seed=9), _bf(N, K, scale=0.05, seed=10)
    ref = dy.double() @ w.double()
    dx = raw.linear_dgrad(dy, w, None, torch.bfloat16)
    _ran_2cta(aux=0, act=0, colsum=0, b_mn=1)
    assert_close(dx, ref, torch.bfloat16, name)
    dx2, sums = raw.linear_dgrad(dy, w, None, torch.bfloat16, colsum=True)
    _ran_2cta(aux=0, colsum=1)
    assert torch.equal(dx2, dx)
    floor = 1e-3 * float(dx.float().abs().max()) * M ** 0.5
    assert rel_err(sums, dx.double().sum(dim=0), floor=floor) < 1e-4


@pytest.mark.parametrize("name,M,N,K", [("c4_fc2", M4, 768, 3072), ("c2_fc2", M2, 384, 1536)], ids=["c4_fc2", "c2_fc2"])
def test_2cta_dgrad_dgelu_aux_colsum(name, M, N, K):
    """fc2 dgrad: dHpre = (dY.W2) * gelu'(pre-activation) with the AUX tile prefetched by TMA, plus the column sums
    (fc1's bias gradient)."""
    from favit_b200 import raw
    dy, w = _bf(M, N, seed=11), _bf(N, K, scale=0.05, seed=12)
    pre = _bf(M, K, scale=1.5, seed=13)
    dx, sums = raw.linear_dgrad(dy, w, pre, torch.bfloat16, colsum=True)
    _ran_2cta(aux=1, act=2, colsum=1)
    p = pre.double().requires_grad_(True)
    torch.nn.functional.gelu(p).backward(dy.double() @ w.double())
    assert_close(dx, p.grad, torch.bfloat16, name, factor=2.0)
    floor = 1e-3 * float(dx.float().abs().max()) * M ** 0.5
    assert rel_err(sums, dx.double().sum(dim=0), floor=floor) < 1e-4


WGRAD = [("c4_qkv", M4, 2304, 768), ("c4_proj", M4, 768, 768), ("c4_fc1", M4, 3072, 768), ("c4_fc2", M4, 768, 3072),
         ("c2_fc1", M2, 1536, 384)]


@pytest.mark.parametrize("name,M,N,K", WGRAD, ids=[w[0] for w in WGRAD])
def test_2cta_wgrad_split_k_reduce_add(name, M, N, K):
    """dW = dY^T.X in fp32 (both operands MN-major, split-K partial tiles combined by TMA reduce-add) + bias gradient;
    then the same call accumulating into an existing gradient."""
    from favit_b200 import _lib as L, raw
    dy, x = _bf(M, N, seed=14), _bf(M, K, seed=15)
    dw, db = raw.linear_wgrad(dy, x, want_bias=True)
    k = _ran_2cta(aux=0, c_fp32=1, reduce=1, a_mn=1, b_mn=1)
    assert " splits=1 " not in k, f"expected a split-K launch at this shape: {k}"
    ref = dy.double().t() @ x.double()
    # fp32 accumulation over 50 432 products per element (tensor-pipe accumulate + reduce-add of the split-K partials):
    # measured 3e-5 of max |dW|; the bar is the north star's fp32 tolerance
    assert rel_err(dw, ref) < 1e-4
    assert rel_err(db, dy.double().sum(dim=0)) < 2e-5
    dw2 = torch.full_like(dw, 0.5)
    rc = L.lib().favit_linear_wgrad(dy.data_ptr(), x.data_ptr(), dw2.data_ptr(), None, M, N, K, N, K, K, L.BF16, 1,
                                    torch.cuda.current_stream().cuda_stream)
    L.check(rc, "favit_linear_wgrad(accumulate)")
    _ran_2cta(reduce=1)
    assert rel_err(dw2, ref + 0.5) < 1e-4


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_2cta_raw_operand_layouts_ragged(a_mn, b_mn):
    """The raw entry point with the CTA-pair kernel requested by name (bn = 512): all four operand layouts, M / N / K not
    multiples of the tile (TMA zero fill on loads, clipping on stores), a forced split-K."""
    from favit_b200 import _lib as L
    M, N, K = 256 * 38 + 72, 256 * 2 + 136, 64 * 9 + 24
    A, B = _bf(M, K, seed=16), _bf(N, K, seed=17)
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    ref = A.double() @ B.double().t()
    st = torch.cuda.current_stream().cuda_stream
    for splits in (1, 3):
        C = torch.zeros(M, N, device="cuda", dtype=torch.float32)
        rc = L.lib().favit_gemm_bf16_raw(As.data_ptr(), a_mn, As.stride(0), Bs.data_ptr(), b_mn, Bs.stride(0),
                                         C.data_ptr(), N, L.F32, M, N, K, 512, splits, st)
        L.check(rc, "gemm_bf16_raw")
        _ran_2cta(splits=splits, a_mn=a_mn, b_mn=b_mn)
        assert rel_err(C, ref) < 1e-5
    Cb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    rc = L.lib().favit_gemm_bf16_raw(As.data_ptr(), a_mn, As.stride(0), Bs.data_ptr(), b_mn, Bs.stride(0), Cb.data_ptr(),
                                     N, L.BF16, M, N, K, 512, 1, st)
    L.check(rc, "gemm_bf16_raw")
    _ran_2cta(c_fp32=0)
    assert_close(Cb, ref, torch.bfloat16, "bf16 C")


def test_2cta_by_name_is_refused_when_not_applicable():
    from favit_b200 import _lib as L
    A, B = _bf(300, 128), _bf(200, 128)
    C = torch.empty(300, 200, device="cuda", dtype=torch.float32)
    rc = L.lib().favit_gemm_bf16_raw(A.data_ptr(), 0, 128, B.data_ptr(), 0, 128, C.data_ptr(), 200, L.F32, 300, 200, 128,
                                     512, 1, torch.cuda.current_stream().cuda_stream)
    assert rc == 2 and b"CTA-pair" in L.lib().favit_last_error()


def test_2cta_repeatable_and_independent_of_previous_tile():
    """The accumulator hand-over between tiles (two TMEM buffers, relaxed 'empty' arrivals): 40 back-to-back launches of
    the persistent kernel give bit-identical bf16 results, and so does a launch whose tiles are visited in a different
    order (a different K changes nothing for the rows compared)."""
    from favit_b200 import ops
    x, w = _bf(M4, 768, seed=18), _bf(2304, 768, scale=0.05, seed=19)
    b = torch.randn(2304, device="cuda")
    first, _ = ops.linear_fwd(x, w, b, None, False, torch.bfloat16, False)
    _ran_2cta()
    for _ in range(40):
        again, _ = ops.linear_fwd(x, w, b, None, False, torch.bfloat16, False)
        assert torch.equal(again, first)
    # the first 25 088 rows alone: other clusters own the tiles, the values must not move
    half, _ = ops.linear_fwd(x[:M4 // 2], w, b, None, False, torch.bfloat16, False)
    _ran_2cta()
    assert torch.equal(half, first[:M4 // 2])


@pytest.fixture
def steal_tiles():
    """Work-stealing tile scheduler for the launches of one test (process-wide switch, put back afterwards)."""
    from favit_b200 import raw
    assert raw.gemm_tile_scheduler() == "static"
    raw.gemm_tile_scheduler("steal")
    yield
    raw.gemm_tile_scheduler("static")


def test_2cta_work_stealing_matches_static_schedule(steal_tiles):
    """Every epilogue of the CTA-pair kernel with tiles handed out by work stealing (what the backward GEMMs of
    data-parallel training run): bit-identical to the statically strided launch where every element has one writer
    (plain bf16, GELU dual store, GELU' AUX), to fp32 summation order where partial tiles are reduce-added (split-K)
    or column sums are accumulated atomically."""
    from favit_b200 import ops, raw
    x, w, b = _bf(M4, 768, seed=40), _bf(3072, 768, scale=0.05, seed=41), torch.randn(3072, device="cuda")
    dy, w2, pre = _bf(M4, 768, seed=42), _bf(768, 3072, scale=0.05, seed=43), _bf(M4, 3072, scale=1.5, seed=44)
    xs = _bf(M4, 2304, seed=45)

    def run():
        y, _ = ops.linear_fwd(x, w, b, None, False, torch.bfloat16, False)
        k0 = _ran_2cta(aux=0, act=0)
        g, gpre = ops.linear_fwd(x, w, b, None, True, torch.bfloat16, True)
        k1 = _ran_2cta(aux=0, act=1)
        dx, sums = raw.linear_dgrad(dy, w2, pre, torch.bfloat16, colsum=True)
        k2 = _ran_2cta(aux=1, act=2, colsum=1)
        dw, db = raw.linear_wgrad(dy, xs, want_bias=True)
        k3 = _ran_2cta(c_fp32=1, reduce=1)
        return (y, g, gpre, dx, sums, dw, db), (k0, k1, k2, k3)

    stolen, kernels = run()
    assert all(" steal=1" in k for k in kernels), kernels
    for _ in range(3):                      # the cursors re-arm themselves: repeat launches on the same sets
        again, _ = run()
        assert all(torch.equal(a, s) for a, s in zip(again[:4], stolen[:4]))
    raw.gemm_tile_scheduler("static")
    static, kernels = run()
    assert all(" steal=0" in k for k in kernels), kernels
    raw.gemm_tile_scheduler("steal")
    for name, a, s in zip(("y", "gelu", "pre", "dx"), stolen[:4], static[:4]):
        assert torch.equal(a, s), name
    assert rel_err(stolen[4], static[4], floor=1e-3 * float(static[4].abs().max())) < 1e-4     # atomic column sums
    assert rel_err(stolen[5], static[5]) < 1e-5                                                  # split-K reduce-add
    assert rel_err(stolen[6], static[6]) < 1e-5


def test_2cta_work_stealing_under_sm_contention(steal_tiles):
    """A launch that finds part of the GPU taken (the NCCL all-reduce of data-parallel training; here: a second
    persistent GEMM on another stream) still computes every tile exactly once: CTA pairs that become resident late find
    their lists worked off by the others, or join in.  Interleaved launches on two streams give bit-identical results to
    the launches run alone, and the cursors re-arm themselves (no memset between launches)."""
    from favit_b200 import ops
    x1, w1 = _bf(M4, 768, seed=30), _bf(2304, 768, scale=0.05, seed=31)
    x2, w2 = _bf(M4 // 2 + 40, 3072, seed=32), _bf(768, 3072, scale=0.03, seed=33)
    b1, b2 = torch.randn(2304, device="cuda"), torch.randn(768, device="cuda")
    ref1, _ = ops.linear_fwd(x1, w1, b1, None, False, torch.bfloat16, False)
    _ran_2cta(steal=1)
    ref2, _ = ops.linear_fwd(x2, w2, b2, None, False, torch.bfloat16, False)
    _ran_2cta(steal=1)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for it in range(6):
        outs1, outs2 = [], []
        with torch.cuda.stream(s1):
            for _ in range(4):
                outs1.append(ops.linear_fwd(x1, w1, b1, None, False, torch.bfloat16, False)[0])
        with torch.cuda.stream(s2):
            for _ in range(6):
                outs2.append(ops.linear_fwd(x2, w2, b2, None, False, torch.bfloat16, False)[0])
        torch.cuda.synchronize()
        assert all(torch.equal(o, ref1) for o in outs1), f"stream 1, round {it}"
        assert all(torch.equal(o, ref2) for o in outs2), f"stream 2, round {it}"
    # and alone again afterwards: the cursors are back at zero
    again, _ = ops.linear_fwd(x1, w1, b1, None, False, torch.bfloat16, False)
    assert torch.equal(again, ref1)


# ---------------------------------------------------------------------------------------------------------------------
# SPPP pooling at the benchmark's batch sizes
# ---------------------------------------------------------------------------------------------------------------------
POOL = [("c2", 256, 224, 16, 16, 384, "sppp_pool_fwd_tma_kernel"), ("c5", 32, 512, 8, 64, 384, "sppp_pool_fwd_sorted_kernel"),
        ("c5_k256", 32, 512, 8, 256, 384, "sppp_pool_fwd_sorted_kernel"), ("c2_vitb", 256, 224, 16, 16, 768, "sppp_pool_fwd_tma_kernel")]


@pytest.mark.parametrize("name,B,S,ps,K,D,kernel", POOL, ids=[p[0] for p in POOL])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
def test_pool_and_assign_at_benchmark_batch(name, B, S, ps, K, D, kernel, dtype):
    from favit_b200 import _lib as L, ops
    from favit_b200.sppp import PatchToSuperpixelMapper
    from favit_b200.synth import voronoi_label_maps
    lm = voronoi_label_maps(B, S, K, seed=21, device="cuda", exact_k=True, patch_size=ps)
    a = PatchToSuperpixelMapper(ps).assign_batch(lm, S, r_cap=K)
    ref = oracle.assign_oracle(lm.cpu().numpy(), ps, S)                   # integers: bit-exact
    assert np.array_equal(a.dom.cpu().numpy(), ref["dom"])
    assert np.array_equal(a.slot.cpu().numpy(), ref["slot"])
    assert np.array_equal(a.num_slots.cpu().numpy(), ref["num_slots"])
    assert np.array_equal(a.counts.cpu().numpy(), np.stack(ref["counts"]))
    assert np.array_equal(a.order.cpu().numpy(), np.stack(ref["order"]))
    assert np.array_equal(a.offsets.cpu().numpy(), np.stack(ref["offsets"]))
    P = (S // ps) ** 2
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(B, P, D, device="cuda", generator=g).to(dtype).requires_grad_(True)
    gout = torch.randn(B, K, D, device="cuda", generator=g)
    out = ops.sppp_pool(x, a.slot, a.counts, a.order, a.offsets, a.num_slots, K)
    k = L.last_kernel()
    assert k.startswith(kernel), k
    if kernel == "sppp_pool_fwd_tma_kernel":        # the persistent ring must be exercised: several items per CTA
        grid, items = [int(t.split("=")[1]) for t in k.split()[1:3]]
        assert items >= 2 * grid, k
    slot64 = a.slot.long()
    cnt = a.counts.double()
    ref_out = torch.zeros(B, K, D, device="cuda", dtype=torch.float64)
    ref_out.scatter_add_(1, slot64[:, :, None].expand(-1, -1, D), x.detach().double())
    ref_out /= cnt[:, :, None]
    assert out.dtype == torch.float32
    assert rel_err(out, ref_out) < 2e-6
    # the same oracle the small-shape tests use, on a sample of images (it loops per image on the host)
    idx = [0, B // 2, B - 1]
    ref_cpu = oracle.pool_mean_batched_oracle(x.detach()[idx].float().cpu(), a.slot[idx].cpu().numpy(), K)
    assert rel_err(out[idx], ref_cpu) < 2e-6
    out.backward(gout)
    dx_ref = torch.gather(gout.double() / cnt[:, :, None], 1, slot64[:, :, None].expand(-1, -1, D))
    assert x.grad.dtype == dtype
    assert rel_err(x.grad, dx_ref) < (1e-6 if dtype == torch.float32 else 8e-3)
    # autograd ran the backward on its own thread (favit_last_kernel is thread-local): the same op called from here
    dx = ops.sppp_pool_bwd(gout, a.slot, a.counts, dtype)
    assert L.last_kernel().startswith(("sppp_pool_bwd_rows_kernel", "sppp_pool_bwd_tile_kernel")), L.last_kernel()
    assert torch.equal(dx, x.grad)
    # size-independent property: pooling a constant field gives the constant, whatever the assignment
    ones = torch.ones(B, P, D, device="cuda", dtype=dtype)
    assert rel_err(ops.sppp_pool_fwd(ones, a.order, a.offsets, a.num_slots, K, torch.float32),
                   torch.ones(B, K, D, device="cuda")) < 1e-6


# ---------------------------------------------------------------------------------------------------------------------
# one full-width training step, replayed from the captured CUDA graph
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_grads(model, x, y, forward):
    sd = {k: v.detach().double().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    loss = torch.nn.functional.cross_entropy(forward(sd), y)
    loss.backward()
    return float(loss), {k: sd[k].grad for k, _ in model.named_parameters()}


def test_vitb_width_graph_step_matches_oracle():
    """VisionTransformerMHLA at ViT-B width (D 768, 12 heads, 197 tokens, window 7), depth 2, 256 images, bf16 autocast,
    through engine.TrainStep(cuda_graph=True): three eager steps, capture, replays.  lr = 0 keeps the parameters fixed,
    so every call must reproduce the oracle's loss and gradients (vit_mhla.py:213-259 via oracle.vit_mhla_forward, fp64).
    The batched latent fold (12-layer path of run_blocks), the CTA-pair GEMMs and the whole-sequence attention kernels
    all run inside the graph."""
    from favit_b200 import _lib as L
    from favit_b200.engine import TrainStep
    from favit_b200.models import VisionTransformerMHLA
    torch.manual_seed(7)
    B = 256
    m = VisionTransformerMHLA(img_size=224, patch_size=16, num_classes=1000, embed_dim=768, depth=2, num_heads=12,
                              window_size=7, use_mhla=True).cuda()
    with torch.no_grad():                      # the reference's pretrained path starts latent_proj at identity; move it
        for blk in m.blocks:                   # off both the identity and the N(0, 0.02) init so that the fold matters
            blk.attn.latent_proj.weight.add_(torch.eye(64, device="cuda") * 0.5)
            blk.attn.latent_proj.bias.normal_(std=0.1)
    g = torch.Generator(device="cuda").manual_seed(3)
    xs = [torch.randn(B, 3, 224, 224, device="cuda", generator=g) for _ in range(2)]
    ys = [torch.randint(0, 1000, (B,), device="cuda", generator=g) for _ in range(2)]
    step = TrainStep(m, lr=0.0, weight_decay=0.0, cuda_graph=True)
    n0 = L.launch_count()
    losses = [float(step(xs[i % 2], ys[i % 2])) for i in range(6)]      # calls 4.. are graph replays
    assert step._graph is not None and L.launch_count() > n0
    i_last = 5 % 2
    ref_loss, ref = _oracle_grads(m, xs[i_last], ys[i_last],
                                  lambda sd: oracle.vit_mhla_forward(xs[i_last].double(), sd, 16, 12, 7))
    assert abs(losses[5] - ref_loss) < 3e-2 * max(1.0, abs(ref_loss)), (losses, ref_loss)
    # replay ~= replay ~= eager on the same batch (the folded proj bias is summed with fp32 atomics: last-bit order effects)
    assert abs(losses[3] - losses[5]) < 1e-4 and abs(losses[1] - losses[5]) < 1e-3
    scale = max(float(r.abs().max()) for r in ref.values())
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        # floor: a gradient that is tiny next to its siblings is compared on the siblings' scale
        assert_close(p.grad, ref[k], torch.bfloat16, f"grad {k}", factor=3.0, floor=1e-3 * scale)


def test_sppp_vits_width_graph_step_matches_oracle():
    """SPPPViTMHLA at the C2 width (ViT-S: D 384, 6 heads, 16 superpixels -> 17 tokens), depth 2, 256 images, through the
    captured graph: the batched assignment, the persistent pooling ring and its backward, the centroid kernel and the
    N = 17 attention kernels, against oracle.sppp_vit_mhla_forward on a sample the per-image oracle loop can afford."""
    from favit_b200.engine import TrainStep
    from favit_b200.models import SPPPViTMHLA
    from favit_b200.synth import voronoi_label_maps
    torch.manual_seed(9)
    B, K = 256, 16
    m = SPPPViTMHLA(img_size=224, patch_size=16, num_classes=1000, embed_dim=384, depth=2, num_heads=6, num_superpixels=K,
                    window_size=7, use_mhla=True, pooling_type="mean").cuda()
    m.validate_slots = False
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(B, 3, 224, 224, device="cuda", generator=g)
    y = torch.randint(0, 1000, (B,), device="cuda", generator=g)
    maps = voronoi_label_maps(B, 224, K, seed=8, device="cuda", exact_k=True, patch_size=16)
    step = TrainStep(m, lr=0.0, weight_decay=0.0, cuda_graph=True)
    losses = [float(step(x, y, maps)) for _ in range(5)]
    assert step._graph is not None and abs(losses[-1] - losses[-2]) < 1e-4
    # (1) graph replay of the full batch == eager step of the full batch (same kernels, same inputs).  After a replay
    # the parameters' .grad are the graph's own tensors; the eager step below replaces them with fresh ones.
    grads_graph = {k: p.grad.clone() for k, p in m.named_parameters()}
    l_eager = float(TrainStep(m, lr=0.0, weight_decay=0.0, cuda_graph=False)(x, y, maps))
    assert abs(losses[-1] - l_eager) < 1e-3
    gscale = max(float(p.grad.abs().max()) for p in m.parameters())
    for k, p in m.named_parameters():
        assert rel_err(grads_graph[k], p.grad, floor=1e-3 * gscale) < 2e-3, k
    # (2) the drop-in against the oracle.  The oracle maps patches per image in a host loop, so it runs on a sample of
    # 24 of the 256 images; the model is per image (no cross-sample statistic), which (1) ties back to the full batch.
    idx = torch.arange(0, B, B // 24)[:24]
    sd32 = {k: v.detach().float().cpu().requires_grad_(True) for k, v in m.state_dict().items()}
    ref_logits = oracle.sppp_vit_mhla_forward(x[idx].cpu(), maps[idx].cpu(), sd32, 16, 6, 7, K)
    ref_loss = torch.nn.functional.cross_entropy(ref_logits, y[idx].cpu())
    ref_loss.backward()
    m.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = m(x[idx], maps[idx])
    loss = torch.nn.functional.cross_entropy(logits.float(), y[idx])
    loss.backward()
    assert_close(logits, ref_logits, torch.bfloat16, "logits", factor=2.0)
    assert abs(float(loss) - float(ref_loss)) < 3e-2 * max(1.0, float(ref_loss))
    scale = max(float(v.grad.abs().max()) for v in sd32.values() if v.grad is not None)
    for k, p in m.named_parameters():
        assert_close(p.grad, sd32[k].grad, torch.bfloat16, f"grad {k}", factor=3.0, floor=1e-3 * scale)
