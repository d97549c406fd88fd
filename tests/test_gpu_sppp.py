"""GPU parity: favit SPPP assignment (bit-exact integers) and segment-mean pooling against the golden fixtures produced
by the reference and against the CPU oracle."""
import numpy as np
import pytest
import torch

import oracle
from util import golden, rel_err

pytestmark = pytest.mark.gpu


def _assign(lm, ps, img, r_cap=None):
    from favit_b200.sppp import PatchToSuperpixelMapper
    return PatchToSuperpixelMapper(ps).assign_batch(torch.from_numpy(lm).cuda(), img, r_cap)


def test_assign_and_pool_match_reference_fixtures():
    from favit_b200.sppp import PatchToSuperpixelMapper, SuperpixelPooling
    g = golden("sppp_maps")
    for name in [str(c) for c in g["cases"]]:
        lm = g[f"{name}_map"]
        ps, img = [int(v) for v in g[f"{name}_cfg"]]
        emb = torch.from_numpy(g[f"{name}_emb"]).cuda()
        a = _assign(lm, ps, img)
        dicts = a.to_dicts()
        mapper, pool = PatchToSuperpixelMapper(ps), SuperpixelPooling("mean")
        for b in range(lm.shape[0]):
            keys, lens, flat = g[f"{name}_{b}_keys"], g[f"{name}_{b}_lens"], g[f"{name}_{b}_flat"]
            d = dicts[b]
            assert list(d.keys()) == keys.tolist(), name
            assert [len(v) for v in d.values()] == lens.tolist(), name
            assert [p for v in d.values() for p in v] == flat.tolist(), name
            R = len(keys)
            assert int(a.num_slots[b]) == R
            assert a.counts[b, :R].cpu().numpy().tolist() == lens.tolist()
            # reference-signature path: map_patches on one image, pool with the dict it returned
            d1 = mapper.map_patches(torch.from_numpy(lm[b]).cuda(), img)
            assert list(d1.items()) == list(d.items())
            pooled = pool.pool(emb[b], d1)
            assert pooled.dtype == torch.float32 and pooled.shape == (R, emb.shape[-1])
            assert torch.allclose(pooled.cpu(), torch.from_numpy(g[f"{name}_{b}_pooled"]), rtol=0, atol=2e-6), name
            # a plain dict (not produced by map_patches) goes through the host-built CSR
            pooled2 = pool.pool(emb[b], dict(d))
            assert torch.equal(pooled2, pooled)


@pytest.mark.parametrize("S,ps,K,B", [(32, 4, 4, 5), (224, 16, 16, 8), (224, 4, 16, 2), (512, 8, 64, 2), (96, 32, 9, 3),
                                      (512, 8, 256, 2), (64, 8, 4, 3), (224, 14, 16, 2)])
def test_assign_bit_exact_vs_oracle(S, ps, K, B):
    from favit_b200.synth import voronoi_label_maps
    lm = voronoi_label_maps(B, S, K, seed=S + ps, device="cpu").numpy()
    ref = oracle.assign_oracle(lm, ps, S)
    a = _assign(lm, ps, S)
    assert np.array_equal(a.dom.cpu().numpy(), ref["dom"])
    assert np.array_equal(a.slot.cpu().numpy(), ref["slot"])
    assert np.array_equal(a.num_slots.cpu().numpy(), ref["num_slots"])
    for b in range(B):
        R = int(ref["num_slots"][b])
        assert np.array_equal(a.counts[b, :R].cpu().numpy(), ref["counts"][b])
        assert np.array_equal(a.slot_label[b, :R].cpu().numpy(), ref["slot_label"][b])
        assert np.array_equal(a.offsets[b, :R + 1].cpu().numpy(), ref["offsets"][b])
        assert np.array_equal(a.order[b].cpu().numpy(), ref["order"][b])


@pytest.mark.parametrize("ps,S", [(8, 48), (16, 64), (32, 64), (16, 50)])
def test_assign_vector_path_adversarial(ps, S):
    """The 16-byte-load dominant kernel (patch 8/16/32, even width) and the shared-memory slot kernel on maps built to
    break them: pixel-level noise from a small alphabet (ties in almost every patch, ids that are negative, huge, or
    first seen in non-sorted order), every pixel distinct (as many slots as patches), and one label everywhere."""
    rng = np.random.default_rng(ps + S)
    alphabet = np.asarray([9, -7, 0, 3, 2 ** 40, 11, 5, -2 ** 50], dtype=np.int64)
    noisy = alphabet[rng.integers(0, len(alphabet), size=(3, S, S))]
    two = np.where(rng.random((2, S, S)) < 0.5, 4, 1).astype(np.int64)         # near-ties between two labels
    g = S // ps
    blocks = rng.permutation(g * g).astype(np.int64).reshape(g, g)              # one distinct label per patch
    per_patch = np.kron(blocks, np.ones((ps, ps), dtype=np.int64))
    per_patch = np.pad(per_patch, ((0, S - g * ps), (0, S - g * ps)))[None]
    const = np.full((1, S, S), 17, dtype=np.int64)
    half = np.zeros((1, S, S), dtype=np.int64)
    half[:, :, 1::2] = -3                                                      # exact half/half tie in every patch
    for lm in (noisy, two, per_patch, const, half):
        ref = oracle.assign_oracle(lm, ps, S)
        a = _assign(lm, ps, S)
        assert np.array_equal(a.dom.cpu().numpy(), ref["dom"])
        assert np.array_equal(a.slot.cpu().numpy(), ref["slot"])
        assert np.array_equal(a.num_slots.cpu().numpy(), ref["num_slots"])
        for b in range(lm.shape[0]):
            R = int(ref["num_slots"][b])
            assert np.array_equal(a.counts[b, :R].cpu().numpy(), ref["counts"][b])
            assert np.array_equal(a.slot_label[b, :R].cpu().numpy(), ref["slot_label"][b])
            assert np.array_equal(a.offsets[b, :R + 1].cpu().numpy(), ref["offsets"][b])
            assert np.array_equal(a.order[b].cpu().numpy(), ref["order"][b])


def test_assign_r_cap_smaller_than_slots():
    """More slots than r_cap: num_slots reports the true R, the first r_cap rows are still exact."""
    from favit_b200.synth import voronoi_label_maps
    lm = voronoi_label_maps(2, 128, 16, seed=5, device="cpu").numpy()
    ref = oracle.assign_oracle(lm, 8, 128)
    a = _assign(lm, 8, 128, r_cap=5)
    assert np.array_equal(a.num_slots.cpu().numpy(), ref["num_slots"])
    assert np.array_equal(a.slot.cpu().numpy(), ref["slot"])
    for b in range(2):
        assert np.array_equal(a.counts[b].cpu().numpy(), ref["counts"][b][:5])
        assert np.array_equal(a.offsets[b].cpu().numpy(), ref["offsets"][b][:6])


def test_assign_adversarial_maps():
    rng = np.random.default_rng(3)
    # random labels per pixel from a small alphabet: many ties, negative and huge ids
    alphabet = np.asarray([-7, 0, 3, 2 ** 40, 11, 5], dtype=np.int64)
    lm = alphabet[rng.integers(0, len(alphabet), size=(4, 24, 24))]
    ref = oracle.assign_oracle(lm, 4, 24)
    a = _assign(lm, 4, 24)
    assert np.array_equal(a.dom.cpu().numpy(), ref["dom"])
    assert np.array_equal(a.slot.cpu().numpy(), ref["slot"])
    assert np.array_equal(a.num_slots.cpu().numpy(), ref["num_slots"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("B,S,ps,K,D", [(4, 32, 4, 4, 24), (8, 224, 16, 16, 384), (2, 128, 8, 16, 100),
                                        (2, 512, 8, 64, 384), (3, 512, 8, 256, 192), (2, 224, 4, 16, 72),
                                        (5, 64, 8, 4, 7), (2, 256, 16, 16, 776)])
def test_pool_fwd_bwd_vs_oracle(B, S, ps, K, D, dtype):
    from favit_b200 import ops
    from favit_b200.synth import voronoi_label_maps
    lm = voronoi_label_maps(B, S, K, seed=11, device="cpu", exact_k=True, patch_size=ps)
    a = _assign(lm.numpy(), ps, S, r_cap=K)
    P = (S // ps) ** 2
    torch.manual_seed(0)
    x = torch.randn(B, P, D).to(dtype)
    g = torch.randn(B, K, D)
    slot = a.slot.cpu().numpy()
    ref = oracle.pool_mean_batched_oracle(x.float(), slot, K)
    xc = x.cuda().requires_grad_(True)
    out = ops.sppp_pool(xc, a.slot, a.counts, a.order, a.offsets, a.num_slots, K)
    assert out.dtype == torch.float32
    assert rel_err(out, ref) < 2e-6
    out.backward(g.cuda())
    cnt = torch.from_numpy(np.stack([np.bincount(slot[b], minlength=K) for b in range(B)])).double()
    s64 = torch.from_numpy(slot.astype(np.int64))
    dx_ref = torch.gather(g.double() / cnt[:, :, None], 1, s64[:, :, None].expand(-1, -1, D))
    assert xc.grad.dtype == dtype
    assert rel_err(xc.grad, dx_ref) < (1e-6 if dtype == torch.float32 else 8e-3)


def test_unequal_slot_counts_raise():
    from favit_b200.sppp import PatchToSuperpixelMapper, SuperpixelPooling
    lm = torch.zeros(2, 16, 16, dtype=torch.int64)
    lm[1, :, 8:] = 1                      # image 0 has one superpixel, image 1 has two
    a = PatchToSuperpixelMapper(4).assign_batch(lm.cuda(), 16)
    x = torch.randn(2, 16, 8, device="cuda")
    with pytest.raises(RuntimeError):
        SuperpixelPooling("mean").pool_batch(x, a, 2)
    with pytest.raises(ValueError):
        SuperpixelPooling("median").pool(x[0], {0: [0]})


@pytest.mark.parametrize("kind", ["max", "attention"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("B,S,ps,K,D", [(3, 32, 4, 4, 24), (2, 224, 16, 16, 384), (2, 128, 8, 16, 100)])
def test_pool_variants_fwd_bwd_vs_oracle(B, S, ps, K, D, dtype, kind):
    """SuperpixelPooling('max' / 'attention') (sppp.py:178-184, 211-216): forward and the gradient w.r.t. the patch
    embeddings against the reference's slot-by-slot torch ops in fp64 (autograd)."""
    from favit_b200.sppp import SuperpixelPooling
    from favit_b200.synth import voronoi_label_maps
    lm = voronoi_label_maps(B, S, K, seed=13, device="cpu", exact_k=True, patch_size=ps)
    a = _assign(lm.numpy(), ps, S, r_cap=K)
    P = (S // ps) ** 2
    torch.manual_seed(1)
    x = (torch.randn(B, P, D) * (0.3 if kind == "attention" else 1.0)).to(dtype)   # row sums are softmax logits
    g = torch.randn(B, K, D)
    slot = a.slot.cpu().numpy()
    x64 = x.double().requires_grad_(True)
    ref = oracle.pool_variant_batched_oracle(x64, slot, K, kind)
    (ref * g.double()).sum().backward()
    xc = x.cuda().requires_grad_(True)
    out = SuperpixelPooling(kind).pool_batch(xc, a, K)
    assert out.dtype == torch.float32 and out.shape == (B, K, D)
    assert rel_err(out, ref.detach()) < 1e-5   # inputs are identical: fp32 math only
    out.backward(g.cuda())
    assert xc.grad.dtype == dtype
    assert rel_err(xc.grad, x64.grad) < (5e-6 if dtype == torch.float32 else 8e-3)
    # reference signature: one image, dict from map_patches
    from favit_b200.sppp import PatchToSuperpixelMapper
    d = PatchToSuperpixelMapper(ps).map_patches(lm[0].cuda(), S)
    one = SuperpixelPooling(kind).pool(x[0].cuda(), d)
    assert torch.allclose(one, out[0].detach(), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("B,H,W,K", [(3, 32, 32, 4), (2, 224, 224, 16), (2, 37, 53, 9), (1, 512, 512, 64), (2, 64, 48, 300)])
def test_centroids_match_reference_loop(B, H, W, K):
    """favit::sppp_centroids against the reference's per-image / per-label loop (sppp_mhla.py:236-262) restated in fp64:
    labels outside [0, K) are ignored, labels without pixels give (0.5, 0.5)."""
    from favit_b200 import ops
    rng = np.random.default_rng(B * 1000 + H + K)
    lm = rng.integers(-2, K + 3, size=(B, H, W)).astype(np.int64)
    lm[:, : H // 2, : W // 2] = rng.integers(0, max(K // 2, 1), size=(B, 1, 1))   # large uniform regions (long runs)
    if K > 3:
        lm[lm == K - 1] = 0                                                       # a label that never occurs
    ref = np.full((B, K, 2), 0.5)
    ys, xs = np.meshgrid(np.arange(H) / H, np.arange(W) / W, indexing="ij")
    for b in range(B):
        for s in range(K):
            m = lm[b] == s
            if m.any():
                ref[b, s, 0] = xs[m].mean()
                ref[b, s, 1] = ys[m].mean()
    got = ops.sppp_centroids(torch.from_numpy(lm).cuda(), K)
    assert got.dtype == torch.float32 and got.shape == (B, K, 2)
    assert np.abs(got.cpu().numpy() - ref).max() < 1e-6


def test_pool_follows_a_dict_edited_after_map_patches():
    """The reference API hands out a plain dict: callers may drop / merge / grow superpixels before pooling.  The cached
    device CSR must not survive such an edit (dict mutation, or in-place edits of the lists inside)."""
    from favit_b200.sppp import PatchToSuperpixelMapper, SuperpixelPooling
    from favit_b200.synth import voronoi_label_maps
    lm = voronoi_label_maps(1, 64, 4, seed=2, device="cuda", exact_k=True, patch_size=8)[0]
    x = torch.randn(64, 24, device="cuda")
    mapper, pool = PatchToSuperpixelMapper(8), SuperpixelPooling("mean")
    ref = lambda d: torch.stack([x[v].mean(dim=0) for v in d.values()])
    d = mapper.map_patches(lm, 64)
    assert d.assignment is not None
    assert torch.allclose(pool.pool(x, d), ref(d), atol=1e-6)
    k0, k1 = list(d)[:2]
    d[k0] = d[k0] + d.pop(k1)                       # merge two superpixels: R drops by one
    assert d.assignment is None
    out = pool.pool(x, d)
    assert out.shape == (len(d), 24) and torch.allclose(out, ref(d), atol=1e-6)
    d2 = mapper.map_patches(lm, 64)
    d2[k1].pop()                                    # in-place edit of a list: the dict itself does not notice
    out2 = pool.pool(x, d2)
    assert torch.allclose(out2, ref(d2), atol=1e-6)
    with pytest.raises(IndexError):
        pool.pool(x, {0: [0, 1], 1: [64]})          # patch id out of range: the reference's fancy indexing raises too


@pytest.mark.parametrize("B,S,ps,K,C", [(3, 32, 4, 4, 3), (4, 224, 16, 16, 3), (2, 512, 8, 64, 3), (2, 64, 8, 16, 1),
                                        (256, 224, 16, 16, 3)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_pool_pixels_is_patchify_then_segment_mean(B, S, ps, K, C, dtype):
    """favit_sppp_pool_pixels against the reference's two steps on raw pixels: the einops rearrangement of
    models/vit.py:38-39 followed by the per-superpixel mean of models/sppp.py:209-210 (fp64)."""
    from favit_b200 import ops
    from favit_b200.sppp import PatchToSuperpixelMapper
    from favit_b200.synth import voronoi_label_maps
    lm = voronoi_label_maps(B, S, K, seed=31, device="cuda", exact_k=True, patch_size=ps)
    a = PatchToSuperpixelMapper(ps).assign_batch(lm, S, r_cap=K)
    g = torch.Generator(device="cuda").manual_seed(7)
    img = torch.randn(B, C, S, S, device="cuda", generator=g)
    out = ops.sppp_pool_pixels(img, a.order, a.offsets, a.num_slots, ps, K, dtype)
    gs = S // ps
    patches = img.double().reshape(B, C, gs, ps, gs, ps).permute(0, 2, 4, 3, 5, 1).reshape(B, gs * gs, ps * ps * C)
    F = ps * ps * C
    ref = torch.zeros(B, K, F, device="cuda", dtype=torch.float64)
    ref.scatter_add_(1, a.slot.long()[:, :, None].expand(-1, -1, F), patches)
    ref /= a.counts.double()[:, :, None]
    assert out.dtype == dtype and out.shape == (B, K, F)
    assert rel_err(out, ref) < (2e-6 if dtype == torch.float32 else 8e-3)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_fused_patch_pool_model_path_equals_two_step_path(mode):
    """SPPPViTMHLA with the algebraic fusion (pool pixels, project R rows) against the same model with the reference's
    two steps (project P rows, pool): logits and every gradient, incl. the patch-embedding weight and bias."""
    from favit_b200.models import SPPPViTMHLA
    from favit_b200.synth import voronoi_label_maps
    torch.manual_seed(2)
    m = SPPPViTMHLA(img_size=64, patch_size=8, num_classes=7, embed_dim=128, depth=2, num_heads=2, num_superpixels=16,
                    window_size=7, use_mhla=True).cuda()
    x = torch.randn(6, 3, 64, 64, device="cuda")
    y = torch.randint(0, 7, (6,), device="cuda")
    maps = voronoi_label_maps(6, 64, 16, seed=5, device="cuda", exact_k=True, patch_size=8)
    res = {}
    for fused in (True, False):
        m.fuse_patch_pool = fused
        m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            logits = m(x, maps)
        torch.nn.functional.cross_entropy(logits.float(), y).backward()
        res[fused] = (logits.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()})
    tol = 1e-4 if mode == "fp32" else 3e-2
    assert rel_err(res[True][0], res[False][0]) < tol
    scale = max(float(v.abs().max()) for v in res[False][1].values())
    for k in res[False][1]:
        assert rel_err(res[True][1][k], res[False][1][k], floor=1e-2 * scale) < 3 * tol, k


@pytest.mark.parametrize("B,S,ps,K", [(5, 224, 16, 16), (3, 512, 8, 64), (2, 96, 32, 9), (3, 64, 8, 4), (2, 100, 8, 9),
                                      (2, 224, 14, 16), (256, 224, 16, 16)])
def test_assign_with_centroids_in_one_pass_equals_the_two_kernels(B, S, ps, K):
    """favit_sppp_assign_centroids (dominant labels + centroid sums from ONE pass over the label map) against the two
    separate entry points: every integer output identical, centroids identical (both are exact integer sums divided
    once), for tiling patch sizes (fused pass), a ragged image (100 % 8 != 0) and patch 14 (separate passes), and with
    labels outside [0, K) and absent labels."""
    from favit_b200 import _lib as L, ops
    from favit_b200.synth import voronoi_label_maps
    lm = voronoi_label_maps(B, S, K, seed=S + K, device="cuda")
    lm[0][lm[0] == 1] = K + 5            # a label >= K: dominates patches, ignored by the centroids; label 1 is now absent
    lm[B - 1][: S // 3, : S // 2] = -3     # negative labels
    got = ops.sppp_assign_centroids(lm, ps, S, K, K)
    fused = "centroids=1" in L.last_kernel()
    assert fused == (S % ps == 0 and ps in (8, 16, 32)), L.last_kernel()
    ref = ops.sppp_assign(lm, ps, S, K)
    for a, b in zip(got[:7], ref):
        assert torch.equal(a, b)
    cen = ops.sppp_centroids(lm, K)
    assert torch.equal(got[7], cen)
    assert torch.allclose(got[7][0, 1], torch.tensor([0.5, 0.5], device="cuda"))
    import oracle
    idx = [0, B - 1]
    assert torch.allclose(got[7][idx].cpu(), oracle.superpixel_centroids(lm[idx].cpu(), K), atol=1e-5)


def test_pool_kernels_order_correctly_behind_and_ahead_of_their_neighbours():
    """The pool kernels are launched with programmatic stream serialization (they may become resident while the previous
    kernel still runs, and let the next one do the same): every global access must sit behind griddepcontrol.wait.
    Producer -> pool -> overwrite-the-input chains, back to back, many times; any premature read or write shows up as a
    mismatch with the result computed from a private copy."""
    from favit_b200 import ops
    from favit_b200.sppp import PatchToSuperpixelMapper
    from favit_b200.synth import voronoi_label_maps
    B, S, ps, K, D = 64, 224, 16, 16, 384
    lm = voronoi_label_maps(B, S, K, seed=9, device="cuda", exact_k=True, patch_size=ps)
    a = PatchToSuperpixelMapper(ps).assign_batch(lm, S, r_cap=K)
    x = torch.randn(B, 196, D, device="cuda").to(torch.bfloat16)
    g = torch.randn(B, K, D, device="cuda")
    outs, refs, dxs, dref = [], [], [], []
    for i in range(40):
        x.mul_(1.03).add_(0.01)                            # producer kernels right in front of the pool forward
        keep = x.clone()
        outs.append(ops.sppp_pool_fwd(x, a.order, a.offsets, a.num_slots, K, torch.float32))
        out2 = ops.sppp_pool_fwd(x, a.order, a.offsets, a.num_slots, K, torch.float32)   # pool directly behind pool
        x.zero_()                                          # a consumer that destroys the input right behind it
        x.copy_(keep)
        refs.append((keep, out2))
        g.mul_(0.99)
        gk = g.clone()
        dxs.append(ops.sppp_pool_bwd(g, a.slot, a.counts, torch.bfloat16))
        g.add_(1.0)                                        # overwrite the backward's input right behind it
        dref.append(gk)
        g.copy_(gk)
    torch.cuda.synchronize()
    for out, (keep, out2), dx, gk in zip(outs, refs, dxs, dref):
        ref = ops.sppp_pool_fwd(keep, a.order, a.offsets, a.num_slots, K, torch.float32)
        assert torch.equal(out, ref) and torch.equal(out2, ref)
        assert torch.equal(dx, ops.sppp_pool_bwd(gk, a.slot, a.counts, torch.bfloat16))


@pytest.mark.parametrize("B,R,D", [(3, 16, 384), (2, 64, 192), (256, 16, 384), (1, 4, 6)])
def test_embed_tokens_is_cls_concat_plus_dynamic_positional_encoding(B, R, D):
    """favit_sppp_embed_tokens against the reference's ops (sppp_mhla.py:302-310 + DynamicPositionalEncoding.forward,
    sppp.py:271-299), forward and the gradients of the pooled tokens and the class token."""
    from favit_b200 import ops
    from favit_b200.models import DynamicPositionalEncoding
    torch.manual_seed(B + R)
    pooled = torch.randn(B, R, D, device="cuda", requires_grad=True)
    cls = torch.randn(1, 1, D, device="cuda", requires_grad=True)
    cen = torch.rand(B, R, 2, device="cuda")
    g = torch.randn(B, R + 1, D, device="cuda")
    out = ops.sppp_embed_tokens(pooled, cls, cen)
    out.backward(g)
    p2, c2 = pooled.detach().clone().requires_grad_(True), cls.detach().clone().requires_grad_(True)
    ref = DynamicPositionalEncoding(D)(torch.cat((c2.expand(B, -1, -1), p2), dim=1), cen)
    ref.backward(g)
    assert torch.allclose(out, ref, rtol=0, atol=5e-6)
    assert torch.allclose(pooled.grad, p2.grad) and torch.allclose(cls.grad, c2.grad, rtol=1e-5, atol=1e-5)
