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
