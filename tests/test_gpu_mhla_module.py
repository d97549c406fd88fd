"""GPU parity of the drop-in MultiHeadLatentAttention / MHLATransformerBlock against fixtures produced by EXECUTING THE
REFERENCE (tests/golden/make_golden.py): output, input gradient and every parameter gradient."""
import pytest
import torch

from util import assert_close, golden

pytestmark = pytest.mark.gpu

PARAMS = ["qkv.weight", "qkv.bias", "proj.weight", "proj.bias", "latent_proj.weight", "latent_proj.bias"]


def _build(g, name, D, H, W):
    from favit_b200.mhla import MultiHeadLatentAttention
    m = MultiHeadLatentAttention(embed_dim=D, num_heads=H, window_size=W)
    sd = {k: torch.from_numpy(g[f"{name}_p_{k}"]).float() for k in PARAMS}
    m.load_state_dict(sd, strict=True)            # same state_dict keys as the reference
    return m.cuda()


@pytest.mark.parametrize("mode", ["fp32", "bf16_autocast", "bf16_input"])
def test_module_matches_reference(mode):
    g = golden("mhla_module")
    dtype = torch.float32 if mode == "fp32" else torch.bfloat16
    for name in [str(c) for c in g["cases"]]:
        B, N, D, H, W, use_mask = [int(v) for v in g[f"{name}_cfg"]]
        if D // H not in (16, 32, 64, 128):
            continue
        m = _build(g, name, D, H, W)
        x = torch.from_numpy(g[f"{name}_x"]).float().cuda().requires_grad_(True)
        gy = torch.from_numpy(g[f"{name}_g"]).float().cuda()
        mask = torch.from_numpy(g[f"{name}_mask"]).float().cuda() if use_mask else None
        if mode == "fp32":
            y = m(x, mask)
        elif mode == "bf16_autocast":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = m(x, mask)
        else:
            y = m(x.to(torch.bfloat16), mask)
        assert y.dtype == dtype, (mode, y.dtype)          # bf16 out under autocast, like nn.Linear
        y.backward(gy.to(y.dtype))
        assert_close(y, torch.from_numpy(g[f"{name}_y"]), dtype, f"{name} y")
        assert_close(x.grad, torch.from_numpy(g[f"{name}_dx"]), dtype, f"{name} dx", factor=2.0)
        for k in PARAMS:
            p = dict(m.named_parameters())[k]
            assert p.grad is not None and p.grad.dtype == torch.float32, k
            assert_close(p.grad, torch.from_numpy(g[f"{name}_dp_{k}"]), dtype, f"{name} d{k}", factor=2.0)


def test_block_matches_reference():
    from favit_b200.mhla import MHLATransformerBlock
    g = golden("mhla_module")
    blk = MHLATransformerBlock(embed_dim=32, num_heads=2, window_size=7, mlp_ratio=2.0)
    sd = {k[len("block_sd_"):]: torch.from_numpy(g[k]).float() for k in g.files if k.startswith("block_sd_")}
    blk.load_state_dict(sd, strict=True)
    blk = blk.cuda()
    y = blk(torch.from_numpy(g["block_x"]).float().cuda())
    assert_close(y, torch.from_numpy(g["block_y"]), torch.float32, "block y")


def test_window_table_matches_reference_fixture():
    from favit_b200.mhla import MultiHeadLatentAttention
    g = golden("mhla_index")
    for key in g.files:
        if not key.startswith("idx_"):
            continue
        n, w = key[4:].split("_")
        m = MultiHeadLatentAttention(embed_dim=8, num_heads=1, window_size=int(w[1:]))
        assert torch.equal(m._get_window_indices(int(n[1:])), torch.from_numpy(g[key])), key
    with pytest.raises(RuntimeError):
        MultiHeadLatentAttention(8, 1, 4)._get_window_indices(6)


def test_constructor_contract():
    from favit_b200.mhla import MultiHeadLatentAttention
    with pytest.raises(AssertionError):
        MultiHeadLatentAttention(embed_dim=30, num_heads=4)
    m = MultiHeadLatentAttention(64, 1, 4).cuda()
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 9, 64, device="cuda"))
    assert list(dict(m.named_parameters())) == PARAMS


def test_unsupported_configurations_are_named_up_front():
    """INTEGRATION.md 'Limits': the favit path raises one clear error at the module boundary."""
    import pytest
    import torch
    from favit_b200.mhla import MultiHeadLatentAttention
    m = MultiHeadLatentAttention(embed_dim=96, num_heads=2).cuda()            # head_dim 48
    with pytest.raises(ValueError, match="head_dim"):
        m(torch.randn(1, 9, 96, device="cuda"))
    m = MultiHeadLatentAttention(embed_dim=64, num_heads=2).cuda()
    with pytest.raises(TypeError, match="float32 or bfloat16"):
        m.half()(torch.randn(1, 9, 64, device="cuda").half())
    with pytest.raises(TypeError, match="float32 or bfloat16"):
        with torch.autocast("cuda", dtype=torch.float16):
            m.float()(torch.randn(1, 9, 64, device="cuda"))
    with pytest.raises(RuntimeError, match="CUDA"):
        MultiHeadLatentAttention(embed_dim=64, num_heads=2)(torch.randn(1, 9, 64))
    # attention-probability dropout must refuse CUDA-graph capture (its seed is a host value)
    m = MultiHeadLatentAttention(embed_dim=64, num_heads=2, dropout=0.2).cuda().train()
    x = torch.randn(2, 9, 64, device="cuda")
    m(x)
    g = torch.cuda.CUDAGraph()
    with pytest.raises(RuntimeError, match="CUDA graph"):
        with torch.cuda.graph(g):
            m(x)
