"""GPU parity of the drop-in models (callers of the hot path) against fixtures produced by executing the reference
models (tests/golden/make_golden.py): logits, loss and every parameter gradient."""
import pytest
import torch

from util import assert_close, golden

pytestmark = pytest.mark.gpu


def _load_sd(g, prefix):
    return {k[len(prefix) + 4:]: torch.from_numpy(g[k]).float() for k in g.files if k.startswith(prefix + "_sd_")}


def _check_grads(model, g, prefix, dtype, factor):
    for k, p in model.named_parameters():
        ref = torch.from_numpy(g[f"{prefix}_grad_{k}"])
        assert p.grad is not None, k
        assert_close(p.grad, ref, dtype, f"{prefix} grad {k}", factor=factor)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_vit_mhla_matches_reference(mode):
    from favit_b200.models import VisionTransformerMHLA
    g = golden("models")
    m = VisionTransformerMHLA(img_size=16, patch_size=4, num_classes=5, embed_dim=32, depth=2, num_heads=2,
                              window_size=3, use_mhla=True)
    m.load_state_dict(_load_sd(g, "vit"), strict=True)
    m = m.cuda()
    x = torch.from_numpy(g["vit_x"]).float().cuda()
    labels = torch.from_numpy(g["vit_labels"]).cuda()
    dtype = torch.float32 if mode == "fp32" else torch.bfloat16
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
        y = m(x)
        loss = torch.nn.functional.cross_entropy(y.float(), labels)
    loss.backward()
    assert_close(y, torch.from_numpy(g["vit_y"]), dtype, "logits", factor=2.0)
    assert abs(loss.item() - float(g["vit_loss"])) < (2e-4 if mode == "fp32" else 3e-2)
    _check_grads(m, g, "vit", dtype, factor=3.0)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_sppp_vit_mhla_matches_reference(mode):
    from favit_b200.models import SPPPViTMHLA
    g = golden("models")
    m = SPPPViTMHLA(img_size=32, patch_size=8, num_classes=5, embed_dim=32, depth=2, num_heads=2, num_superpixels=4,
                    window_size=3, use_mhla=True, pooling_type="mean")
    m.load_state_dict(_load_sd(g, "sppp"), strict=True)
    m = m.cuda()
    x = torch.from_numpy(g["sppp_x"]).cuda()
    seg = torch.from_numpy(g["sppp_maps"]).cuda()
    labels = torch.from_numpy(g["sppp_labels"]).cuda()
    dtype = torch.float32 if mode == "fp32" else torch.bfloat16
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
        y = m(x, seg)
        loss = torch.nn.functional.cross_entropy(y.float(), labels)
    loss.backward()
    assert_close(y, torch.from_numpy(g["sppp_y"]), dtype, "logits", factor=2.0)
    assert abs(loss.item() - float(g["sppp_loss"])) < (2e-4 if mode == "fp32" else 3e-2)
    _check_grads(m, g, "sppp", dtype, factor=3.0)
    # the reference hook still works: replace model.segmentation.segment instead of passing the maps
    m.segmentation.segment = lambda img: seg
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
        y2 = m(x)
    assert torch.equal(y2, y)


def test_centroids_match_oracle():
    import oracle
    from favit_b200.models import SPPPViTMHLA
    from favit_b200.synth import voronoi_label_maps
    m = SPPPViTMHLA(img_size=64, patch_size=8, num_classes=3, embed_dim=32, depth=1, num_heads=1, num_superpixels=16,
                    use_mhla=True)
    seg = voronoi_label_maps(3, 64, 16, seed=4, device="cpu")
    seg[2][seg[2] == 5] = 99          # a label outside 0..K-1 and an absent label -> (0.5, 0.5)
    got = m._calculate_superpixel_centroids(seg.cuda()).cpu()
    assert torch.allclose(got, oracle.superpixel_centroids(seg, 16), atol=1e-5)


@pytest.mark.parametrize("B,C,S,ps", [(3, 3, 32, 4), (2, 3, 224, 16), (2, 1, 64, 8), (256, 3, 224, 16)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_patchify_is_the_einops_rearrangement(B, C, S, ps, dtype):
    """favit_patchify == einops 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' (models/vit.py:38-39) followed by the cast."""
    from favit_b200 import ops
    x = torch.randn(B, C, S, S, device="cuda")
    got = ops.patchify(x, ps, dtype)
    g = S // ps
    ref = x.reshape(B, C, g, ps, g, ps).permute(0, 2, 4, 3, 5, 1).reshape(B, g * g, ps * ps * C).to(dtype)
    assert got.dtype == dtype and torch.equal(got, ref)


def test_patch_embedding_and_head_run_on_favit_gemms():
    """PatchEmbedding.forward and the classification head go through favit::linear (no cuBLAS launch in the step):
    outputs and weight gradients against nn.Linear on the rearranged patches."""
    from favit_b200 import _lib as L
    from favit_b200.models import PatchEmbedding, favit_linear
    torch.manual_seed(0)
    pe = PatchEmbedding(img_size=64, patch_size=8, in_channels=3, embed_dim=128).cuda()
    x = torch.randn(4, 3, 64, 64, device="cuda")
    n0 = L.launch_count()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = pe(x)
    assert L.launch_count() > n0 and y.dtype == torch.bfloat16 and y.shape == (4, 64, 128)
    y.float().square().sum().backward()
    lin = pe.projection[1]
    gw = lin.weight.grad.clone()
    lin.weight.grad = None
    patches = x.reshape(4, 3, 8, 8, 8, 8).permute(0, 2, 4, 3, 5, 1).reshape(4, 64, 192)
    ref = torch.nn.functional.linear(patches.double(), lin.weight.double(), lin.bias.double())
    assert_close(y, ref, torch.bfloat16, "patch embedding")
    (ref.float().square().sum()).backward()
    assert_close(gw, lin.weight.grad, torch.bfloat16, "patch embedding dW", factor=2.0)
    head = torch.nn.Linear(128, 1000).cuda()
    t = torch.randn(4, 128, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = favit_linear(head, t)
    assert out.dtype == torch.bfloat16
    assert_close(out, torch.nn.functional.linear(t.double(), head.weight.double(), head.bias.double()), torch.bfloat16, "head")
    assert favit_linear(torch.nn.Linear(128, 10).cuda(), t).shape == (4, 10)      # 10 classes: not a multiple of 8 -> nn.Linear
