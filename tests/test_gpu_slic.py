"""GPU superpixel segmentation (favit_slic_segment, SURVEY.md §8f-3: the step models/sppp.py:26-74 does with
skimage.segmentation.slic on the CPU) against its CPU restatement oracle/slic_oracle.py — bit-exact labels — and
structural properties; then the reference flow `model(x)` with no label maps given, SLIC running on the device."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _blocks(B, S, g, noise, seed, C=3):
    """g x g constant-colour blocks + noise: an image whose superpixels are known."""
    rng = np.random.default_rng(seed)
    cols = rng.uniform(-2, 2, size=(B, C, g, g)).astype(np.float32)
    img = np.repeat(np.repeat(cols, S // g, axis=2), S // g, axis=3)
    return (img + noise * rng.standard_normal(img.shape).astype(np.float32)).astype(np.float32)


@pytest.mark.parametrize("B,C,H,W,K,comp,sigma,iters", [(2, 3, 64, 64, 16, 0.1, 1.0, 10), (1, 3, 224, 224, 16, 0.1, 1.0, 10),
                                                        (2, 1, 48, 80, 12, 0.5, 0.0, 3), (1, 3, 100, 60, 9, 10.0, 2.0, 5),
                                                        (3, 3, 32, 32, 4, 0.1, 1.0, 0)])
def test_slic_labels_bit_exact_vs_oracle(B, C, H, W, K, comp, sigma, iters):
    from favit_b200 import ops
    rng = np.random.default_rng(H * W + K)
    img = rng.standard_normal((B, C, H, W)).astype(np.float32)
    img[:, :, : H // 2] += 1.5                         # some structure besides the noise
    got = ops.slic_segment(torch.from_numpy(img).cuda(), K, comp, sigma, iters).cpu().numpy()
    ref = oracle.slic_oracle(img, K, comp, sigma, iters)
    assert got.dtype == np.int64 and got.shape == (B, H, W)
    assert np.array_equal(got, ref), f"{(got != ref).mean():.2e} of the labels differ"
    gy, gx = oracle.slic_grid(H, W, K)
    assert got.min() >= 0 and got.max() < gy * gx
    # a pixel can only join one of the 3 x 3 grid-neighbour clusters of its own cell
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    cy = np.minimum((ys * gy) // H, gy - 1)
    cx = np.minimum((xs * gx) // W, gx - 1)
    assert (np.abs(got // gx - cy) <= 1).all() and (np.abs(got % gx - cx) <= 1).all()


def test_slic_recovers_blocks_and_is_deterministic():
    from favit_b200 import ops
    img = torch.from_numpy(_blocks(4, 64, 4, 0.05, seed=1)).cuda()
    a = ops.slic_segment(img, 16, 0.1, 1.0, 10)
    b = ops.slic_segment(img, 16, 0.1, 1.0, 10)
    assert torch.equal(a, b)                           # fixed-point centre sums: no run-to-run variation
    truth = (torch.arange(64).view(-1, 1) // 16 * 4 + torch.arange(64).view(1, -1) // 16).cuda()
    assert float((a == truth).float().mean()) > 0.95  # block interiors agree; only the blurred borders may move


def test_model_forward_segments_on_the_device_when_no_maps_are_given():
    """The reference's own call `model(x)`: `self.segmentation.segment(x)` (sppp_mhla.py:278) now runs on the GPU."""
    from favit_b200 import _lib as L
    from favit_b200.models import SPPPViTMHLA
    torch.manual_seed(0)
    m = SPPPViTMHLA(img_size=64, patch_size=8, num_classes=5, embed_dim=64, depth=1, num_heads=1, num_superpixels=16,
                    window_size=7, use_mhla=True).cuda()
    x = torch.from_numpy(_blocks(3, 64, 4, 0.05, seed=2)).cuda()
    maps = m.segmentation.segment(x)
    assert maps.dtype == torch.int64 and maps.shape == (3, 64, 64) and maps.is_cuda
    y1 = m(x)                                         # segments internally
    y2 = m(x, maps)                                   # same maps passed explicitly
    assert torch.equal(y1, y2) and y1.shape == (3, 5)
    one = m.segmentation.segment(x[0])                # unbatched call of the reference API
    assert torch.equal(one, maps[0])
