"""CPU: the C-ABI library builds/loads and exports every symbol include/favit.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "favit.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(favit_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import favit_b200
    from favit_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    h = _lib.lib()
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/favit.h but not exported"
    assert sorted(_lib.exported_symbols()) == names, "ctypes signature table and header disagree"
    assert h.favit_version() >= 100
    assert isinstance(h.favit_launch_count(), int)


def test_gemm_tile_scheduler_switch_is_host_state():
    """favit_set_gemm_tile_scheduler: 0 / 1 set the mode, anything else only queries; the Python wrapper names the modes
    and refuses unknown ones.  No device needed: the mode is read when a GEMM is launched."""
    from favit_b200 import _lib, raw
    h = _lib.lib()
    assert h.favit_set_gemm_tile_scheduler(-1) == 0            # default: static striding
    assert h.favit_set_gemm_tile_scheduler(1) == 1
    assert h.favit_set_gemm_tile_scheduler(7) == 1             # out of range: query only
    assert raw.gemm_tile_scheduler() == "steal"
    assert raw.gemm_tile_scheduler("static") == "static"
    assert h.favit_set_gemm_tile_scheduler(-1) == 0
    with pytest.raises(ValueError, match="tile scheduler"):
        raw.gemm_tile_scheduler("dynamic")


def test_ops_fail_loudly_without_cuda():
    import torch
    from favit_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.linear(torch.randn(4, 8), torch.randn(8, 8), None)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.mhla_attn(torch.randn(1, 4, 3, 1, 16), 3, None)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.sppp_assign(torch.zeros(1, 8, 8, dtype=torch.int64), 4, 8, 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "focused-attention-vit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f


def test_host_side_shapes_via_fake_tensors():
    """register_fake kernels give the right shapes/dtypes without touching a device."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    from favit_b200 import ops
    with FakeTensorMode():
        qkv = torch.empty(2, 17, 3, 6, 64, dtype=torch.bfloat16, device="cuda")
        out, lse = ops.mhla_attn(qkv, 7, None)
        assert out.shape == (2, 17, 384) and out.dtype == torch.bfloat16
        assert lse.shape == (2, 6, 17) and lse.dtype == torch.float32
        y = ops.linear(torch.empty(34, 384, dtype=torch.bfloat16, device="cuda"),
                       torch.empty(1152, 384, device="cuda"), torch.empty(1152, device="cuda"))
        assert y.shape == (34, 1152) and y.dtype == torch.bfloat16
        r = ops.sppp_assign(torch.empty(2, 224, 224, dtype=torch.int64, device="cuda"), 16, 224, 16)
        assert [t.shape for t in r] == [(2, 196), (2, 196), (2,), (2, 16), (2, 16), (2, 17), (2, 196)]
        out, lse = ops.mhla_attn(qkv, 7, None, 0.1, 1234)          # attention-probability dropout arguments
        assert out.shape == (2, 17, 384) and lse.shape == (2, 6, 17)
        c = ops.sppp_centroids(torch.empty(2, 224, 224, dtype=torch.int64, device="cuda"), 16)
        assert c.shape == (2, 16, 2) and c.dtype == torch.float32


def test_dropout_keep_mask_is_a_fair_coin_and_seed_dependent():
    """The oracle side of the counter-based dropout mask (the CUDA side is compared with it bit for bit on the GPU)."""
    import numpy as np
    import oracle
    for p in (0.1, 0.5):
        m = oracle.dropout_keep_mask(4, 6, 65, 7, p, seed=99)
        assert m.shape == (4, 6, 65, 7) and abs(m.mean() - (1 - p)) < 0.02
        assert abs(m[..., 0].mean() - m[..., 6].mean()) < 0.05        # no slot bias
    a, b = oracle.dropout_keep_mask(1, 1, 50, 7, 0.5, 1), oracle.dropout_keep_mask(1, 1, 50, 7, 0.5, 2)
    assert (a != b).mean() > 0.3
    assert np.array_equal(a, oracle.dropout_keep_mask(1, 1, 50, 7, 0.5, 1))


def test_oracle_gather_core_equals_closed_form():
    """The two CPU formulations of the attention core (reference order of operations vs banded softmax with integer
    multiplicities) agree, with and without a mask, for N < W, N = 1 and wide windows."""
    import torch
    import oracle
    torch.manual_seed(1)
    for (N, W) in [(10, 7), (3, 7), (1, 7), (40, 15), (17, 31), (6, 1)]:
        q, k, v = [torch.randn(2, 2, N, 16, dtype=torch.float64) for _ in range(3)]
        mask = (torch.rand(2, N, N) > 0.3).double()
        mask[:, torch.arange(N), torch.arange(N)] = 1
        # rows whose whole window is masked are NaN in both; keep the diagonal and the edge keys visible
        mask[:, :, 0] = 1
        mask[:, :, N - 1] = 1
        for m in (None, mask):
            a = oracle.mhla_attn_core_gather(q, k, v, W, m)
            b, _ = oracle.mhla_attn_core_closed_form(q, k, v, W, m)
            assert torch.allclose(a, b, rtol=1e-10, atol=1e-12), (N, W, m is None)
