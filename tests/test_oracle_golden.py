"""CPU: pin the oracle against fixtures produced by executing the reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle.mhla_oracle import window_indices, window_multiplicity

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(G, name + ".npz"), allow_pickle=False)


def test_window_tables_match_reference():
    g = _load("mhla_index")
    n_checked = 0
    for key in g.files:
        if not key.startswith("idx_"):
            continue
        n, w = key[4:].split("_")
        n, w = int(n[1:]), int(w[1:])
        got = window_indices(n, w)
        assert got.dtype == np.int64
        assert np.array_equal(got, g[key]), key
        # multiplicity closed form is consistent with the table
        m = window_multiplicity(n, w)
        ref_m = np.zeros((n, n), dtype=np.int64)
        for i in range(n):
            for j in g[key][i]:
                ref_m[i, j] += 1
        assert np.array_equal(m, ref_m)
        n_checked += 1
    assert n_checked >= 12
    assert int(g["even_window_raises"][0]) == 1
    with pytest.raises(RuntimeError):
        window_indices(6, 4)


def _params(g, name, dtype=torch.float64):
    def t(k):
        return torch.from_numpy(g[f"{name}_p_{k}"]).to(dtype).requires_grad_(True)
    return dict(qkv_w=t("qkv.weight"), qkv_b=t("qkv.bias"), proj_w=t("proj.weight"), proj_b=t("proj.bias"),
                lat_w=t("latent_proj.weight"), lat_b=t("latent_proj.bias"))


GRAD_KEYS = {"qkv_w": "qkv.weight", "qkv_b": "qkv.bias", "proj_w": "proj.weight", "proj_b": "proj.bias",
             "lat_w": "latent_proj.weight", "lat_b": "latent_proj.bias"}


@pytest.mark.parametrize("form", ["gather", "closed"])
def test_mhla_module_matches_reference_fp64(form):
    g = _load("mhla_module")
    fn = oracle.mhla_forward_gather if form == "gather" else oracle.mhla_forward_closed_form
    for name in [str(c) for c in g["cases"]]:
        B, N, D, H, W, use_mask = [int(v) for v in g[f"{name}_cfg"]]
        p = _params(g, name)
        x = torch.from_numpy(g[f"{name}_x"]).requires_grad_(True)
        mask = torch.from_numpy(g[f"{name}_mask"]) if use_mask else None
        y = fn(x, num_heads=H, window_size=W, attention_mask=mask, **p)
        (y * torch.from_numpy(g[f"{name}_g"])).sum().backward()
        assert torch.allclose(y, torch.from_numpy(g[f"{name}_y"]), rtol=1e-10, atol=1e-12), name
        assert torch.allclose(x.grad, torch.from_numpy(g[f"{name}_dx"]), rtol=1e-9, atol=1e-12), name
        for k, ref_k in GRAD_KEYS.items():
            ref = torch.from_numpy(g[f"{name}_dp_{ref_k}"])
            assert torch.allclose(p[k].grad, ref, rtol=1e-9, atol=1e-11), (name, k)


def test_sppp_assign_and_pool_match_reference():
    g = _load("sppp_maps")
    for name in [str(c) for c in g["cases"]]:
        lm = g[f"{name}_map"]
        ps, img = [int(v) for v in g[f"{name}_cfg"]]
        emb = torch.from_numpy(g[f"{name}_emb"])
        res = oracle.assign_oracle(lm, ps, img)
        for b in range(lm.shape[0]):
            keys, lens, flat = g[f"{name}_{b}_keys"], g[f"{name}_{b}_lens"], g[f"{name}_{b}_flat"]
            # literal loop
            d = oracle.map_patches_oracle(lm[b], ps, img)
            assert list(d.keys()) == keys.tolist(), name
            assert [len(v) for v in d.values()] == lens.tolist()
            assert [p for v in d.values() for p in v] == flat.tolist()
            # vectorised arrays, bit-exact
            assert int(res["num_slots"][b]) == len(keys)
            assert np.array_equal(res["slot_label"][b], keys)
            assert np.array_equal(res["counts"][b], lens.astype(np.int32))
            assert np.array_equal(res["order"][b], flat.astype(np.int32))
            slot_ref = np.empty(flat.shape[0], dtype=np.int32)
            off = 0
            for r, n in enumerate(lens):
                slot_ref[flat[off:off + n]] = r
                off += n
            assert np.array_equal(res["slot"][b], slot_ref)
            assert np.array_equal(res["dom"][b], keys[slot_ref])
            pooled = oracle.pool_mean_oracle(emb[b], d)
            assert pooled.dtype == torch.float32
            assert torch.allclose(pooled, torch.from_numpy(g[f"{name}_{b}_pooled"]), rtol=0, atol=1e-6)
            batched = oracle.pool_mean_batched_oracle(emb[b:b + 1], res["slot"][b:b + 1], len(keys))
            assert torch.allclose(batched[0].float(), pooled, atol=1e-6)


def _model_sd(g, prefix, dtype):
    sd = {}
    for k in g.files:
        if k.startswith(prefix + "_sd_"):
            sd[k[len(prefix) + 4:]] = torch.from_numpy(g[k]).to(dtype).requires_grad_(True)
    return sd


def test_vit_mhla_model_matches_reference_fp64():
    g = _load("models")
    sd = _model_sd(g, "vit", torch.float64)
    y = oracle.vit_mhla_forward(torch.from_numpy(g["vit_x"]), sd, patch_size=4, num_heads=2, window=3)
    assert torch.allclose(y, torch.from_numpy(g["vit_y"]), rtol=1e-9, atol=1e-11)
    loss = torch.nn.functional.cross_entropy(y, torch.from_numpy(g["vit_labels"]))
    assert abs(loss.item() - float(g["vit_loss"])) < 1e-10
    loss.backward()
    for k, p in sd.items():
        ref = torch.from_numpy(g[f"vit_grad_{k}"])
        assert torch.allclose(p.grad, ref, rtol=1e-8, atol=1e-11), k


def test_sppp_vit_mhla_model_matches_reference_fp32():
    g = _load("models")
    sd = _model_sd(g, "sppp", torch.float32)
    seg = torch.from_numpy(g["sppp_maps"])
    y = oracle.sppp_vit_mhla_forward(torch.from_numpy(g["sppp_x"]), seg, sd, patch_size=8, num_heads=2, window=3,
                                     num_superpixels=4)
    assert torch.allclose(y, torch.from_numpy(g["sppp_y"]), rtol=1e-4, atol=1e-5)
    loss = torch.nn.functional.cross_entropy(y, torch.from_numpy(g["sppp_labels"]))
    assert abs(loss.item() - float(g["sppp_loss"])) < 1e-5
    loss.backward()
    for k, p in sd.items():
        ref = torch.from_numpy(g[f"sppp_grad_{k}"])
        scale = max(ref.abs().max().item(), 1e-6)
        assert (p.grad - ref).abs().max().item() <= 2e-4 * scale + 1e-7, k
