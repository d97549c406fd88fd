"""CPU: the oracle against the REFERENCE ITSELF, executed live on randomised inputs — beyond the committed fixtures.

Runs only where /root/reference is mounted (the build container); on the GPU box, which has no reference, every test
here is skipped and the fixtures of tests/golden/ (outputs of the same reference, tests/test_oracle_golden.py) carry
the pin.  The reference is imported exactly as tests/golden/make_golden.py does it (scikit-image's `slic` stubbed:
label maps are inputs of the path).  Nothing of the product is involved: this file only strengthens the
checker that the CUDA path is compared with.

Covered, each over many seeds / shapes (the fixtures hold 10 module cases and 9 label-map cases):
  * `_get_window_indices` for every (N, W) with N <= 48 and odd W <= 33, plus even W with N <= W (models/mhla.py:46-83);
  * `MultiHeadLatentAttention` forward, input gradient and all six parameter gradients in fp64, with and without masks,
    N < W, N = 1, against BOTH oracle formulations (gather: op for op; closed form + latent fold: what the kernels
    implement) (models/mhla.py:85-161);
  * `map_patches` dictionaries (keys, order, patch lists) on random label maps with ties, negative and huge ids, ragged
    image sizes (models/sppp.py:91-128), and the batched arrays of `assign_oracle` derived from them;
  * `SuperpixelPooling('mean' | 'max' | 'attention').pool`, 2-D and 3-D inputs (models/sppp.py:153-223);
  * `_calculate_superpixel_centroids` and the centroid branch of `DynamicPositionalEncoding`
    (models/sppp_mhla.py:226-262, models/sppp.py:271-299);
  * whole models, random small configurations: logits and every parameter gradient.
"""
import os
import sys
import zlib

import numpy as np
import pytest
import torch

import oracle

REF = os.environ.get("FAVIT_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")),
                                reason="the reference is not mounted here; tests/golden/ carries the pin")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    return make_golden._import_reference()


def _maxerr(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max())


def test_window_tables_exhaustive(ref):
    checked = 0
    for W in list(range(1, 34, 2)) + [2, 4, 6, 8]:
        m = ref["MHLA"](embed_dim=8, num_heads=1, window_size=W)
        for N in range(1, 49):
            if W % 2 == 0 and N > W:
                with pytest.raises(RuntimeError):
                    m._get_window_indices(N)
                continue
            want = m._get_window_indices(N).numpy()
            got = oracle.window_indices(N, W)
            assert np.array_equal(got, want), (N, W)
            # the multiplicity form the kernels use is the same table, counted
            mult = oracle.window_multiplicity(N, W)
            assert np.array_equal(mult, np.stack([np.bincount(r, minlength=N) for r in want])), (N, W)
            checked += 1
    assert checked > 800


MODULE_CASES = [(seed, B, N, H, hd, W, masked)
                for seed, (B, N, H, hd, W, masked) in enumerate([
                    (2, 23, 2, 8, 7, False), (1, 4, 3, 4, 7, False), (3, 1, 1, 8, 5, False), (2, 9, 2, 16, 9, True),
                    (1, 31, 4, 4, 15, False), (2, 12, 1, 8, 1, False), (1, 40, 2, 8, 31, False), (2, 6, 2, 4, 6, False),
                    (2, 17, 3, 8, 3, True), (1, 66, 1, 16, 33, False), (2, 5, 2, 8, 9, True), (1, 50, 2, 4, 7, True)])]


@pytest.mark.parametrize("seed,B,N,H,hd,W,masked", MODULE_CASES)
def test_mhla_module_forward_and_gradients(ref, seed, B, N, H, hd, W, masked):
    torch.manual_seed(1000 + seed)
    D = H * hd
    mod = ref["MHLA"](embed_dim=D, num_heads=H, window_size=W).double()
    with torch.no_grad():                       # the default init leaves latent_proj far from identity: keep it generic
        for p in mod.parameters():
            p.add_(0.05 * torch.randn_like(p))
    x = torch.randn(B, N, D, dtype=torch.float64, requires_grad=True)
    g = torch.randn(B, N, D, dtype=torch.float64)
    mask = None
    if masked:
        mask = (torch.rand(B, N, N) > 0.35).double()
        mask[:, torch.arange(N), torch.arange(N)] = 1.0            # no fully masked row (NaN rows have their own test)
    y = mod(x, mask)
    (y * g).sum().backward()
    want = {n: p.grad.clone() for n, p in mod.named_parameters()}
    names = ["qkv.weight", "qkv.bias", "proj.weight", "proj.bias", "latent_proj.weight", "latent_proj.bias"]
    for form in (oracle.mhla_forward_gather, oracle.mhla_forward_closed_form):
        ps = [mod.get_parameter(n).detach().clone().requires_grad_(True) for n in names]
        xo = x.detach().clone().requires_grad_(True)
        yo = form(xo, *ps, H, W, mask)
        (yo * g).sum().backward()
        assert _maxerr(yo, y) < 1e-11, form.__name__
        assert _maxerr(xo.grad, x.grad) < 1e-10, form.__name__
        for n, p in zip(names, ps):
            if form is oracle.mhla_forward_closed_form and n == "latent_proj.bias":
                # K-path share is exactly zero; the closed form routes everything through the V path
                assert _maxerr(p.grad, want[n]) < 1e-9, n
            else:
                assert _maxerr(p.grad, want[n]) < 1e-9, (form.__name__, n)


def test_fully_masked_row_gives_nan_like_the_reference(ref):
    torch.manual_seed(5)
    mod = ref["MHLA"](embed_dim=16, num_heads=2, window_size=5).double()
    x = torch.randn(1, 8, 16, dtype=torch.float64)
    mask = torch.ones(1, 8, 8, dtype=torch.float64)
    mask[0, 3, :] = 0
    y = mod(x, mask)
    ps = [p.detach() for p in (mod.qkv.weight, mod.qkv.bias, mod.proj.weight, mod.proj.bias, mod.latent_proj.weight,
                               mod.latent_proj.bias)]
    yo = oracle.mhla_forward_gather(x, *ps, 2, 5, mask)
    assert torch.isnan(y[0, 3]).all() and torch.isnan(yo[0, 3]).all()
    keep = [i for i in range(8) if i != 3]
    assert _maxerr(yo[0, keep], y[0, keep]) < 1e-12


def _random_maps(rng, kind, S, B=2):
    if kind == "voronoi":
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
        import make_golden
        return make_golden.voronoi_maps(rng, B, S, 9, jitter=0.45)
    if kind == "noise":                       # every patch a handful of labels: ties everywhere
        return rng.integers(0, 4, size=(B, S, S)).astype(np.int64)
    if kind == "wild":                        # negative and huge ids, first-seen order far from sorted order
        ids = np.asarray([2 ** 40, -7, 3, 99999, -(2 ** 33), 0, 12], dtype=np.int64)
        blocks = rng.integers(0, len(ids), size=(B, (S + 5) // 6, (S + 5) // 6))
        return ids[np.repeat(np.repeat(blocks, 6, axis=1), 6, axis=2)[:, :S, :S]]
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["voronoi", "noise", "wild"])
@pytest.mark.parametrize("S,ps", [(32, 4), (48, 8), (36, 8), (64, 16), (30, 4)])
def test_map_patches_and_pooling(ref, kind, S, ps):
    rng = np.random.default_rng(zlib.crc32(f"{kind}-{S}-{ps}".encode()))
    lm = _random_maps(rng, kind, S)
    mapper = ref["Mapper"](patch_size=ps)
    g = S // ps
    P = g * g
    arrays = oracle.assign_oracle(lm, ps, S)
    torch.manual_seed(P)
    emb = torch.randn(lm.shape[0], P, 12, dtype=torch.float64)
    for b in range(lm.shape[0]):
        want = mapper.map_patches(torch.from_numpy(lm[b]), S)
        got = oracle.map_patches_oracle(lm[b], ps, S)
        assert list(got.keys()) == list(want.keys()), "slot order = first appearance in raster patch order"
        assert all(got[k] == want[k] for k in want)
        # the batched arrays say the same thing
        assert arrays["num_slots"][b] == len(want)
        assert np.array_equal(arrays["slot_label"][b], np.asarray(list(want.keys()), dtype=np.int64))
        for r, patches in enumerate(want.values()):
            o0, o1 = arrays["offsets"][b][r], arrays["offsets"][b][r + 1]
            assert list(arrays["order"][b][o0:o1]) == patches
            assert all(arrays["slot"][b][p] == r for p in patches)
        R = len(want)
        for kind_p in ("mean", "max", "attention"):
            pool = ref["Pool"](kind_p)
            e = emb[b].clone().requires_grad_(True)
            pooled = pool.pool(e, want)                                    # 2-D branch
            assert pooled.dtype == torch.float32                          # torch.zeros default dtype (sppp.py:198)
            if kind_p == "mean":
                mine = oracle.pool_mean_batched_oracle(emb[b:b + 1], arrays["slot"][b:b + 1], R)[0]
                assert _maxerr(oracle.pool_mean_oracle(emb[b], want), pooled) < 1e-6
            else:
                mine = oracle.pool_variant_batched_oracle(emb[b:b + 1], arrays["slot"][b:b + 1], R, kind_p)[0]
            assert _maxerr(mine, pooled) < 1e-6, kind_p                   # the reference rounds its output to fp32
            # 3-D branch with a batch that shares the dictionary (sppp.py:160-190)
            pooled3 = pool.pool(emb[b:b + 1].expand(2, -1, -1), want)
            assert _maxerr(pooled3[1], pooled) < 1e-6


def test_unknown_pooling_type_raises_value_error(ref):
    pool = ref["Pool"]("median")
    with pytest.raises(ValueError):
        pool.pool(torch.randn(4, 3), {0: [0, 1], 1: [2, 3]})
    with pytest.raises(ValueError):
        oracle.pool_variant_batched_oracle(torch.randn(1, 4, 3), np.asarray([[0, 0, 1, 1]]), 2, "median")


@pytest.mark.parametrize("seed", range(4))
def test_centroids_and_dynamic_positional_encoding(ref, seed):
    rng = np.random.default_rng(seed)
    K, S, D = 9, 24 + 6 * seed, 16
    lm = _random_maps(rng, "voronoi", S, B=2)
    if seed % 2:
        lm[lm == 4] = 3                        # an absent label: the reference falls back to (0.5, 0.5)
    sp = ref["SPPPViT"](img_size=S, patch_size=6, num_classes=3, embed_dim=D, depth=1, num_heads=2, num_superpixels=K,
                        window_size=3, use_mhla=True)
    seg = torch.from_numpy(lm)
    want_c = sp._calculate_superpixel_centroids(seg)
    got_c = oracle.superpixel_centroids(seg, K)
    assert _maxerr(got_c, want_c) < 1e-6
    torch.manual_seed(seed)
    x = torch.randn(2, K + 1, D)
    want = sp.pos_embed(x, want_c)             # dropout 0
    assert _maxerr(oracle.dynamic_positional_encoding(x, got_c), want) < 1e-6


@pytest.mark.parametrize("seed,cfg", list(enumerate([
    dict(img_size=16, patch_size=4, embed_dim=24, depth=1, num_heads=3, window_size=5),
    dict(img_size=24, patch_size=8, embed_dim=32, depth=2, num_heads=2, window_size=7),       # N = 10, short windows
    dict(img_size=20, patch_size=4, embed_dim=16, depth=3, num_heads=1, window_size=3, mlp_ratio=2.0)])))
def test_vit_mhla_model_logits_and_gradients(ref, seed, cfg):
    torch.manual_seed(40 + seed)
    vit = ref["ViT"](num_classes=6, use_mhla=True, **cfg).double()
    x = torch.randn(2, 3, cfg["img_size"], cfg["img_size"], dtype=torch.float64)
    labels = torch.tensor([seed, 5 - seed])
    y = vit(x)
    torch.nn.functional.cross_entropy(y, labels).backward()
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in vit.state_dict().items()}
    yo = oracle.vit_mhla_forward(x, sd, cfg["patch_size"], cfg["num_heads"], cfg["window_size"])
    torch.nn.functional.cross_entropy(yo, labels).backward()
    assert _maxerr(yo, y) < 1e-10
    for k, p in vit.named_parameters():
        assert _maxerr(sd[k].grad, p.grad) < 1e-9, k


@pytest.mark.parametrize("seed", range(3))
def test_sppp_vit_mhla_model_logits_and_gradients(ref, seed):
    torch.manual_seed(60 + seed)
    rng = np.random.default_rng(60 + seed)
    S, ps, K = (32, 8, 4) if seed < 2 else (48, 8, 9)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    lm = make_golden.voronoi_maps(rng, 2, S, K, jitter=0.2)
    sp = ref["SPPPViT"](img_size=S, patch_size=ps, num_classes=4, embed_dim=32, depth=2, num_heads=2, num_superpixels=K,
                        window_size=3 + 2 * seed, use_mhla=True, pooling_type="mean")
    sp.segmentation.segment = lambda img: torch.from_numpy(lm)
    x = torch.randn(2, 3, S, S)
    labels = torch.tensor([1, 3])
    try:
        y = sp(x)
    except RuntimeError:
        pytest.skip("a superpixel dominated no patch in this draw: the reference itself raises (R != num_superpixels)")
    torch.nn.functional.cross_entropy(y, labels).backward()
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in sp.state_dict().items()}
    yo = oracle.sppp_vit_mhla_forward(x, torch.from_numpy(lm), sd, ps, 2, 3 + 2 * seed, K)
    torch.nn.functional.cross_entropy(yo, labels).backward()
    assert _maxerr(yo, y) < 2e-5                                           # fp32 model
    for k, p in sp.named_parameters():
        scale = max(float(p.grad.abs().max()), 1e-3)
        assert _maxerr(sd[k].grad, p.grad) / scale < 2e-4, k
