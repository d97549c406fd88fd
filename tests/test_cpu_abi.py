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
