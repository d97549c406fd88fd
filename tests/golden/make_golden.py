"""Generate the golden fixtures by EXECUTING THE REFERENCE (read-only, /root/reference).

Run once in the build container (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
It writes tests/golden/{mhla_index,mhla_module,sppp_maps,models}.npz.  Nothing here is imported by the
product; the fixtures pin the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.

scikit-image is not installed, and the label maps are *inputs* of the hot path, so
`skimage.segmentation.slic` is stubbed before importing `models.sppp`.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("FAVIT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, REF)
    sk = types.ModuleType("skimage")
    seg = types.ModuleType("skimage.segmentation")
    seg.slic = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("slic is stubbed; label maps are inputs"))
    sk.segmentation = seg
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.segmentation", seg)
    from models.mhla import MultiHeadLatentAttention, MHLATransformerBlock
    from models.sppp import PatchToSuperpixelMapper, SuperpixelPooling
    from models.vit_mhla import VisionTransformerMHLA
    from models.sppp_mhla import SPPPViTMHLA
    return dict(MHLA=MultiHeadLatentAttention, Block=MHLATransformerBlock, Mapper=PatchToSuperpixelMapper,
                Pool=SuperpixelPooling, ViT=VisionTransformerMHLA, SPPPViT=SPPPViTMHLA)


def voronoi_maps(rng, B, S, K, jitter=0.35, labels=None):
    """Jittered-grid Voronoi label maps."""
    k = int(round(K ** 0.5))
    assert k * k == K
    cell = S / k
    cy, cx = np.meshgrid((np.arange(k) + 0.5) * cell, (np.arange(k) + 0.5) * cell, indexing="ij")
    yy, xx = np.meshgrid(np.arange(S) + 0.5, np.arange(S) + 0.5, indexing="ij")
    out = np.empty((B, S, S), dtype=np.int64)
    for b in range(B):
        sy = cy.reshape(-1) + rng.uniform(-jitter, jitter, K) * cell
        sx = cx.reshape(-1) + rng.uniform(-jitter, jitter, K) * cell
        d = (yy[None] - sy[:, None, None]) ** 2 + (xx[None] - sx[:, None, None]) ** 2
        lab = np.argmin(d, axis=0)
        if labels is not None:
            lab = np.asarray(labels, dtype=np.int64)[lab]
        out[b] = lab
    return out


def gen_index(ref, out):
    cases = [(10, 7), (5, 7), (3, 7), (1, 7), (65, 7), (17, 7), (197, 7), (8, 1), (6, 3), (3, 4), (4, 4),
             (20, 15), (17, 31), (40, 9)]
    for n, w in cases:
        m = ref["MHLA"](embed_dim=8, num_heads=1, window_size=w)
        out[f"idx_N{n}_W{w}"] = m._get_window_indices(n).numpy().astype(np.int64)
    # the ragged case must raise RuntimeError (mhla.py:83)
    raised = 0
    try:
        ref["MHLA"](embed_dim=8, num_heads=1, window_size=4)._get_window_indices(6)
    except RuntimeError:
        raised = 1
    out["even_window_raises"] = np.asarray([raised])


def gen_module(ref, out):
    # (name, B, N, D, H, W, mask?)
    cases = [("a", 2, 10, 32, 2, 7, False), ("short", 1, 5, 32, 2, 7, False), ("hd64", 2, 17, 64, 1, 7, False),
             ("w3", 1, 12, 48, 3, 3, False), ("mask", 2, 10, 32, 2, 7, True), ("w1", 1, 6, 16, 1, 1, False),
             ("n1", 2, 1, 32, 2, 7, False), ("w15", 1, 40, 64, 2, 15, False),
             # wide windows, head_dim 64: the tcgen05 / TMEM attention kernels (round 2)
             ("w31", 1, 70, 64, 1, 31, False), ("w63", 2, 130, 128, 2, 63, False)]
    names = []
    for name, B, N, D, H, W, use_mask in cases:
        torch.manual_seed(sum(ord(c) for c in name))
        mod = ref["MHLA"](embed_dim=D, num_heads=H, window_size=W).double()
        x = torch.randn(B, N, D, dtype=torch.float64, requires_grad=True)
        g = torch.randn(B, N, D, dtype=torch.float64)
        mask = None
        if use_mask:
            mask = (torch.rand(B, N, N) > 0.3).to(torch.float64)
            for i in range(N):
                mask[:, i, i] = 1.0            # never fully mask a row
        y = mod(x, mask)
        (y * g).sum().backward()
        out[f"{name}_cfg"] = np.asarray([B, N, D, H, W, int(use_mask)])
        out[f"{name}_x"] = x.detach().numpy()
        out[f"{name}_g"] = g.numpy()
        if mask is not None:
            out[f"{name}_mask"] = mask.numpy()
        for pn, p in mod.named_parameters():
            out[f"{name}_p_{pn}"] = p.detach().numpy()
            out[f"{name}_dp_{pn}"] = p.grad.numpy()
        out[f"{name}_y"] = y.detach().numpy()
        out[f"{name}_dx"] = x.grad.numpy()
        names.append(name)
    out["cases"] = np.asarray(names)
    # a transformer block (mhla.py:164-222)
    torch.manual_seed(7)
    blk = ref["Block"](embed_dim=32, num_heads=2, window_size=7, mlp_ratio=2.0).double()
    x = torch.randn(2, 9, 32, dtype=torch.float64)
    out["block_x"] = x.numpy()
    out["block_y"] = blk(x).detach().numpy()
    for pn, p in blk.state_dict().items():
        out[f"block_sd_{pn}"] = p.numpy()


def gen_sppp(ref, out):
    rng = np.random.default_rng(20261018)
    maps = {}
    maps["vor32"] = (voronoi_maps(rng, 3, 32, 4), 4, 32)
    maps["vor224"] = (voronoi_maps(rng, 2, 224, 16), 16, 224)
    maps["vor64p8"] = (voronoi_maps(rng, 2, 64, 16), 8, 64)
    # first-seen order != sorted order, non-contiguous label ids >= K
    maps["relabel"] = (voronoi_maps(rng, 2, 48, 9, labels=[1000, 7, 50, 3, 99999, 12, 2 ** 40, 5, 4]), 8, 48)
    # exact ties: every patch is half label 9 / half label 2 (left/right) -> smaller id wins
    tie = np.empty((1, 32, 32), dtype=np.int64)
    tie[:, :, :] = 9
    for j in range(0, 32, 8):
        tie[:, :, j + 4:j + 8] = 2
    tie[:, 16:, :] = np.where(tie[:, 16:, :] == 2, 11, 4)       # lower half: 4 vs 11 tie -> 4
    maps["ties"] = (tie, 8, 32)
    maps["single"] = (np.full((1, 16, 16), 5, dtype=np.int64), 4, 16)
    # every pixel a different label inside a patch (all counts 1 -> smallest id wins), negatives too
    distinct = (np.arange(16 * 16, dtype=np.int64).reshape(1, 16, 16) * 37) % 251 - 100
    maps["distinct"] = (distinct, 4, 16)
    # img_size % patch_size != 0: trailing pixels ignored (sppp.py:102)
    maps["ragged"] = (voronoi_maps(rng, 1, 36, 4), 8, 36)
    names = []
    for name, (lm, ps, img) in maps.items():
        mapper = ref["Mapper"](patch_size=ps)
        pool = ref["Pool"]("mean")
        out[f"{name}_map"] = lm
        out[f"{name}_cfg"] = np.asarray([ps, img])
        P = (img // ps) ** 2
        torch.manual_seed(P)
        emb = torch.randn(lm.shape[0], P, 24)
        out[f"{name}_emb"] = emb.numpy()
        for b in range(lm.shape[0]):
            d = mapper.map_patches(torch.from_numpy(lm[b]), img)
            keys = np.asarray(list(d.keys()), dtype=np.int64)
            lens = np.asarray([len(v) for v in d.values()], dtype=np.int64)
            flat = np.asarray([p for v in d.values() for p in v], dtype=np.int64)
            out[f"{name}_{b}_keys"] = keys
            out[f"{name}_{b}_lens"] = lens
            out[f"{name}_{b}_flat"] = flat
            out[f"{name}_{b}_pooled"] = pool.pool(emb[b], d).numpy()
        names.append(name)
    out["cases"] = np.asarray(names)


def gen_models(ref, out):
    torch.manual_seed(11)
    vit = ref["ViT"](img_size=16, patch_size=4, num_classes=5, embed_dim=32, depth=2, num_heads=2,
                     window_size=3, use_mhla=True).double()
    x = torch.randn(3, 3, 16, 16, dtype=torch.float64)
    out["vit_x"] = x.numpy()
    y = vit(x)
    out["vit_y"] = y.detach().numpy()
    labels = torch.tensor([1, 0, 4])
    loss = torch.nn.functional.cross_entropy(y, labels)
    loss.backward()
    out["vit_labels"] = labels.numpy()
    out["vit_loss"] = np.asarray(loss.item())
    for k, v in vit.state_dict().items():
        out[f"vit_sd_{k}"] = v.numpy()
    for k, p in vit.named_parameters():
        out[f"vit_grad_{k}"] = p.grad.numpy()

    torch.manual_seed(12)
    rng = np.random.default_rng(5)
    lm = voronoi_maps(rng, 2, 32, 4)
    sp = ref["SPPPViT"](img_size=32, patch_size=8, num_classes=5, embed_dim=32, depth=2, num_heads=2,
                        num_superpixels=4, window_size=3, use_mhla=True, pooling_type="mean")
    sp.segmentation.segment = lambda img: torch.from_numpy(lm)
    x = torch.randn(2, 3, 32, 32)
    y = sp(x)
    labels = torch.tensor([2, 3])
    loss = torch.nn.functional.cross_entropy(y, labels)
    loss.backward()
    out["sppp_x"] = x.numpy()
    out["sppp_maps"] = lm
    out["sppp_y"] = y.detach().numpy()
    out["sppp_labels"] = labels.numpy()
    out["sppp_loss"] = np.asarray(loss.item())
    for k, v in sp.state_dict().items():
        out[f"sppp_sd_{k}"] = v.numpy()
    for k, p in sp.named_parameters():
        out[f"sppp_grad_{k}"] = p.grad.numpy()


def main():
    ref = _import_reference()
    for fname, fn in [("mhla_index", gen_index), ("mhla_module", gen_module), ("sppp_maps", gen_sppp),
                      ("models", gen_models)]:
        d = {}
        fn(ref, d)
        path = os.path.join(HERE, fname + ".npz")
        np.savez_compressed(path, **d)
        print(f"{path}: {len(d)} arrays, {os.path.getsize(path)/1024:.1f} KiB")


if __name__ == "__main__":
    main()
