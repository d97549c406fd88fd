"""GPU numerics of the tcgen05 bf16 GEMM and the fp32 SIMT GEMM behind favit::linear: forward (+bias, GELU, residual),
dgrad (+GELU'), wgrad + bias gradient, against a plain PyTorch fp64 reference of the same op."""
import pytest
import torch

from util import assert_close, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("bn", [0, 64, 128, 256])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 200, 136), (1000, 768, 384), (50, 2304, 768)])
def test_raw_gemm_all_operand_layouts(a_mn, b_mn, bn, M, N, K):
    from favit_b200 import _lib as L
    torch.manual_seed(M + N + K)
    # MN-major operands need the M / N extent to be a multiple of 8 (TMA row pitch)
    if a_mn:
        M = (M + 7) // 8 * 8
    if b_mn:
        N = (N + 7) // 8 * 8
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    C = torch.empty(M, N, device="cuda", dtype=torch.float32)
    rc = L.lib().favit_gemm_bf16_raw(As.data_ptr(), a_mn, As.stride(0), Bs.data_ptr(), b_mn, Bs.stride(0), C.data_ptr(),
                                     N, L.F32, M, N, K, bn, 1, torch.cuda.current_stream().cuda_stream)
    L.check(rc, "gemm_bf16_raw")
    ref = A.double() @ B.double().t()
    assert rel_err(C, ref) < 1e-5          # bf16 products are exact in fp32; only the summation order differs


@pytest.mark.parametrize("splits", [2, 5])
def test_raw_gemm_split_k(splits):
    from favit_b200 import _lib as L
    M, N, K = 256, 384, 64 * 23
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    C = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    rc = L.lib().favit_gemm_bf16_raw(A.data_ptr(), 0, K, B.data_ptr(), 0, K, C.data_ptr(), N, L.F32, M, N, K, 128, splits,
                                     torch.cuda.current_stream().cuda_stream)
    L.check(rc, "gemm_bf16_raw")
    assert rel_err(C, A.double() @ B.double().t()) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("M,K,N", [(4160, 192, 576), (197 * 3, 768, 2304), (17 * 16, 384, 1152), (65, 64, 40)])
def test_linear_fwd_bwd(dtype, M, K, N):
    from favit_b200 import ops
    torch.manual_seed(K)
    x = torch.randn(M, K, device="cuda").to(dtype).requires_grad_(True)
    w = (torch.randn(N, K, device="cuda") * 0.05).requires_grad_(True)       # fp32 master weight
    b = torch.randn(N, device="cuda").requires_grad_(True)
    gy = torch.randn(M, N, device="cuda").to(dtype)
    y = ops.linear(x, w, b)
    y.backward(gy)
    assert y.dtype == dtype and w.grad.dtype == torch.float32 and b.grad.dtype == torch.float32
    xr = x.detach().double().requires_grad_(True)
    wr = w.detach().to(dtype).double().requires_grad_(True)     # the kernel sees the weight in the compute dtype
    br = b.detach().double().requires_grad_(True)
    yr = torch.nn.functional.linear(xr, wr, br)
    yr.backward(gy.double())
    assert_close(y, yr, dtype, "y")
    assert_close(x.grad, xr.grad, dtype, "dx")
    assert_close(w.grad, wr.grad, torch.float32 if dtype == torch.float32 else dtype, "dw")
    assert_close(b.grad, br.grad, torch.float32 if dtype == torch.float32 else dtype, "db")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("M,K,N", [(520, 128, 392), (4200, 256, 648)], ids=["small", "cta_pair"])
def test_linear_epilogues(dtype, M, K, N):
    from favit_b200 import ops
    x = torch.randn(M, K, device="cuda").to(dtype)
    w = (torch.randn(N, K, device="cuda") * 0.1).to(dtype)
    b = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    # bias + GELU, saving the pre-activation
    y, pre = ops.linear_fwd(x, w, b, None, True, dtype, True)
    pre_ref = x.double() @ w.double().t() + b.double()
    assert_close(pre, pre_ref, dtype, "preact")
    assert_close(y, torch.nn.functional.gelu(pre_ref), dtype, "gelu")
    # bias + fp32 residual, fp32 output (the residual stream of a bf16 block)
    y2, _ = ops.linear_fwd(x, w, b, res, False, torch.float32, False)
    assert y2.dtype == torch.float32
    assert_close(y2, pre_ref + res.double(), dtype, "residual")
    # dgrad with GELU'
    dy = torch.randn(M, N, device="cuda").to(dtype)
    pre_k = torch.randn(M, K, device="cuda").to(dtype)
    dx = ops.linear_dgrad(dy, w, pre_k, dtype)
    pk = pre_k.double().requires_grad_(True)
    torch.nn.functional.gelu(pk).backward(dy.double() @ w.double())
    assert_close(dx, pk.grad, dtype, "dgelu", factor=2.0)


@pytest.mark.parametrize("M,N,K", [(4200, 648, 256), (300, 200, 136)], ids=["cta_pair_fused", "small_separate_pass"])
def test_dgrad_column_sums(M, N, K):
    """dx_colsum of favit_linear_dgrad == column sums of the dX it stored (fused in the CTA-pair epilogue for large
    shapes, a separate reduction otherwise)."""
    from favit_b200 import raw
    torch.manual_seed(3)
    dy = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
    pre = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    for preact in (None, pre):
        dx, sums = raw.linear_dgrad(dy, w, preact, torch.bfloat16, colsum=True)
        ref = dx.double().sum(dim=0)
        assert rel_err(sums, ref, floor=1e-3 * float(dx.float().abs().max()) * M ** 0.5) < 1e-4


def test_attention_backward_column_sums():
    from favit_b200 import raw
    B, N, H, hd, W = 3, 65, 3, 64, 7
    qkv = torch.randn(B * N, 3 * H * hd, device="cuda").to(torch.bfloat16)
    do = torch.randn(B * N, H * hd, device="cuda").to(torch.bfloat16)
    o, lse = raw.attn_fwd(qkv, B, N, H, hd, W)
    dqkv, sums = raw.attn_bwd(qkv, o, lse, do, B, N, H, hd, W)
    ref = dqkv.double().sum(dim=0)
    assert rel_err(sums, ref, floor=1e-3 * float(dqkv.float().abs().max()) * (B * N) ** 0.5) < 1e-4
