"""CPU checks of the oracle pieces added for the round-2 kernels (their CUDA counterparts are compared with them bit for
bit in the -m gpu suite): the MLP dropout keep-mask generator and the SLIC restatement."""
import numpy as np

import oracle


def test_mlp_dropout_mask_statistics_and_offsets():
    for p in (0.1, 0.5):
        k, inv = oracle.mlp_dropout_keep_mask(512, 770, p, seed=1234, layer=0, site=1)
        assert k.shape == (512, 770) and abs(k.mean() - (1 - p)) < 5e-3
        assert abs(inv - 1 / (1 - p)) < 1e-3 * inv
        assert abs(k[:, 0::4].mean() - k[:, 3::4].mean()) < 1e-2            # the four 16-bit lanes of a hash are unbiased
    a, _ = oracle.mlp_dropout_keep_mask(64, 64, 0.5, 7, 0, 1)
    for other in ((8, 0, 1), (7, 1, 1), (7, 0, 2)):                           # seed, layer, site all change the mask
        b, _ = oracle.mlp_dropout_keep_mask(64, 64, 0.5, *other)
        assert (a != b).mean() > 0.3
    c, _ = oracle.mlp_dropout_keep_mask(64, 64, 0.5, 7, 0, 1)
    assert np.array_equal(a, c)
    none, inv = oracle.mlp_dropout_keep_mask(8, 8, 0.0, 7, 0, 1)
    assert none.all() and inv == 1.0


def test_slic_oracle_recovers_blocks_and_respects_the_grid():
    rng = np.random.default_rng(0)
    cols = rng.uniform(-2, 2, size=(2, 3, 4, 4)).astype(np.float32)
    img = np.repeat(np.repeat(cols, 16, axis=2), 16, axis=3) + 0.05 * rng.standard_normal((2, 3, 64, 64)).astype(np.float32)
    lab = oracle.slic_oracle(img, 16, compactness=0.1, sigma=1.0, iters=10)
    truth = (np.arange(64)[:, None] // 16) * 4 + np.arange(64)[None, :] // 16
    assert lab.dtype == np.int64 and lab.shape == (2, 64, 64)
    assert (lab == truth).mean() > 0.97
    assert oracle.slic_grid(224, 224, 16) == (4, 4) and oracle.slic_grid(100, 300, 24) == (3, 8)
    # zero iterations = nearest initial centre in the joint colour / position metric; labels stay inside the grid
    lab0 = oracle.slic_oracle(img, 9, compactness=10.0, sigma=0.0, iters=0)
    assert lab0.min() >= 0 and lab0.max() < 9
