"""CPU oracle for Superpixel Patch Pooling.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates `/root/reference/models/sppp.py`:
  * PatchToSuperpixelMapper.map_patches  -> sppp.py:91-128  (dominant label per patch: most frequent,
    ties -> smallest label id because torch.unique sorts and argmax returns the first maximum,
    sppp.py:117-120; dict insertion order = order in which dominant labels first appear in raster
    patch order, sppp.py:123-126)
  * SuperpixelPooling.pool ('mean', 2-D branch) -> sppp.py:192-223 (output is fp32 zeros, row i is
    the mean of the patches of the i-th dict entry)
  * per-image loop + torch.stack in the model forwards -> sppp_mhla.py:283-300

`map_patches_oracle` is the literal loop (small cases); `assign_oracle` is a vectorised numpy
restatement that emits the arrays the CUDA path emits (dominant label, slot id, counts, CSR order).
Integer outputs must match the CUDA path bit for bit.

Parity pin: tests/golden/sppp_*.npz, produced by running the reference (tests/golden/make_golden.py).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch


def map_patches_oracle(segmentation_map: np.ndarray, patch_size: int, img_size: int) -> Dict[int, List[int]]:
    """Literal restatement of sppp.py:102-128 on a [H, W] integer map."""
    seg = np.asarray(segmentation_map)
    g = img_size // patch_size
    out: Dict[int, List[int]] = {}
    for i in range(g):
        for j in range(g):
            patch = seg[i * patch_size:(i + 1) * patch_size, j * patch_size:(j + 1) * patch_size]
            labels, counts = np.unique(patch, return_counts=True)      # sorted ascending
            dominant = int(labels[int(np.argmax(counts))])              # first maximum -> smallest label
            out.setdefault(dominant, []).append(i * g + j)
    return out


def assign_oracle(label_maps: np.ndarray, patch_size: int, img_size: int):
    """Batched arrays equivalent to calling map_patches per image.

    Returns dict with
      dom        [B, P] int64   dominant label of each patch
      slot       [B, P] int32   pooled row index of each patch (first-seen order)
      num_slots  [B]    int32   R per image
      counts     list of int32 arrays, counts[b][r] = patches in slot r
      slot_label list of int64 arrays, label owning slot r
      order      list of int32 arrays, patch ids grouped by slot, ascending inside a slot
      offsets    list of int32 arrays, CSR offsets into order (len R+1)
    """
    lm = np.asarray(label_maps)
    if lm.ndim == 2:
        lm = lm[None]
    B = lm.shape[0]
    g = img_size // patch_size
    P = g * g
    ps = patch_size
    tiles = lm[:, :g * ps, :g * ps].reshape(B, g, ps, g, ps).transpose(0, 1, 3, 2, 4).reshape(B, P, ps * ps)
    tiles = np.sort(tiles, axis=-1)
    dom = np.empty((B, P), dtype=np.int64)
    # mode with smallest-label tie-break on the sorted pixels: run lengths
    n = ps * ps
    for b in range(B):
        t = tiles[b]                                                   # [P, n] sorted ascending
        # occurrences of each pixel's label inside its patch; first maximum along the sorted
        # axis is the most frequent label with the smallest id
        chunk = max(1, (1 << 24) // (n * n))
        for p0 in range(0, P, chunk):
            tt = t[p0:p0 + chunk]
            occ = (tt[:, :, None] == tt[:, None, :]).sum(axis=-1)
            k = np.argmax(occ, axis=-1)
            dom[b, p0:p0 + chunk] = np.take_along_axis(tt, k[:, None], axis=1)[:, 0]
    slot = np.empty((B, P), dtype=np.int32)
    num_slots = np.empty((B,), dtype=np.int32)
    counts, slot_label, order, offsets = [], [], [], []
    for b in range(B):
        seen: Dict[int, int] = {}
        for p in range(P):
            lab = int(dom[b, p])
            if lab not in seen:
                seen[lab] = len(seen)
            slot[b, p] = seen[lab]
        R = len(seen)
        num_slots[b] = R
        counts.append(np.bincount(slot[b], minlength=R).astype(np.int32))
        lab_arr = np.empty(R, dtype=np.int64)
        for lab, r in seen.items():
            lab_arr[r] = lab
        slot_label.append(lab_arr)
        order.append(np.argsort(slot[b], kind="stable").astype(np.int32))
        offsets.append(np.concatenate([[0], np.cumsum(counts[-1])]).astype(np.int32))
    return dict(dom=dom, slot=slot, num_slots=num_slots, counts=counts, slot_label=slot_label,
                order=order, offsets=offsets)


def pool_mean_oracle(patch_embeddings: torch.Tensor, superpixel_to_patches: Dict[int, List[int]]) -> torch.Tensor:
    """sppp.py:192-223 (2-D branch, 'mean'): fp32 zeros [R, D]; row i = mean over the i-th entry."""
    D = patch_embeddings.shape[-1]
    out = torch.zeros(len(superpixel_to_patches), D)
    for i, (_, patches) in enumerate(superpixel_to_patches.items()):
        if not patches:
            continue
        out[i, :] = torch.mean(patch_embeddings[patches, :], dim=0)
    return out


def pool_mean_batched_oracle(x: torch.Tensor, slot: np.ndarray, num_slots: int) -> torch.Tensor:
    """Per-image loop + stack of sppp_mhla.py:283-300 written as an index_add (fp64 accumulate)."""
    B, P, D = x.shape
    out = torch.zeros(B, num_slots, D, dtype=torch.float64)
    s = torch.from_numpy(np.asarray(slot).astype(np.int64))
    for b in range(B):
        out[b].index_add_(0, s[b], x[b].to(torch.float64))
        cnt = torch.bincount(s[b], minlength=num_slots).clamp_min(1).to(torch.float64)
        out[b] /= cnt[:, None]
    return out


def pool_variant_batched_oracle(x: torch.Tensor, slot: np.ndarray, num_slots: int, pooling_type: str) -> torch.Tensor:
    """'max' / 'attention' pooling (models/sppp.py:178-184, 211-216) for a batch, slot by slot in the reference's order of
    operations (differentiable torch ops; pass an fp64 leaf to get reference gradients from autograd)."""
    B, P, D = x.shape
    rows = []
    for b in range(B):
        out = []
        for r in range(num_slots):
            idx = np.nonzero(np.asarray(slot[b]) == r)[0]
            if len(idx) == 0:
                out.append(torch.zeros(D, dtype=x.dtype))
                continue
            e = x[b, torch.from_numpy(idx)]
            if pooling_type == "max":
                out.append(torch.max(e, dim=0)[0])
            elif pooling_type == "attention":
                w = torch.softmax(torch.sum(e, dim=-1), dim=-1)
                out.append(torch.sum(e * w.unsqueeze(-1), dim=0))
            else:
                raise ValueError(f"Unsupported pooling type: {pooling_type}")
        rows.append(torch.stack(out))
    return torch.stack(rows)
