"""Drop-in `PatchToSuperpixelMapper` / `SuperpixelPooling` backed by the favit sm_100a kernels.

Reference: /root/reference/models/sppp.py:77-223.  The reference API is per image and dict based
(`map_patches(seg[H,W], img_size) -> {label: [patch ids]}`, `pool(x[N,D], dict) -> [R,D]`); it is kept, and a batched
device-resident path (`assign_batch` / `pool_batch`) replaces the per-image Python loop + `torch.stack` of
models/sppp_mhla.py:283-300 with one launch per batch.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import ops


@dataclass
class SuperpixelAssignment:
    """Device-resident result of the patch -> superpixel assignment for a batch (all bit-exact with the dicts the
    reference builds): see include/favit.h:favit_sppp_assign."""
    dom: torch.Tensor         # [B,P] int64  dominant label of each patch
    slot: torch.Tensor        # [B,P] int32  pooled row of each patch (first-seen order of its dominant label)
    num_slots: torch.Tensor   # [B]   int32  R per image
    counts: torch.Tensor      # [B,r_cap] int32
    slot_label: torch.Tensor  # [B,r_cap] int64
    offsets: torch.Tensor     # [B,r_cap+1] int32
    order: torch.Tensor       # [B,P] int32  patch ids grouped by slot, ascending inside a slot
    r_cap: int

    def to_dicts(self) -> List[Dict[int, List[int]]]:
        """Materialise the reference's per-image dicts (one D2H copy; insertion order = slot order)."""
        ns = self.num_slots.cpu().tolist()
        lab = self.slot_label.cpu().tolist()
        off = self.offsets.cpu().tolist()
        order = self.order.cpu().tolist()
        out = []
        for b, R in enumerate(ns):
            if R > self.r_cap:
                raise RuntimeError(f"image {b} has {R} superpixel slots, more than r_cap={self.r_cap}")
            d = AssignmentDict()
            for r in range(R):
                d[int(lab[b][r])] = order[b][off[b][r]:off[b][r + 1]]
            out.append(d)
        return out


class AssignmentDict(dict):
    """The dict `map_patches` returns; remembers the device arrays it was built from so `pool` need not rebuild
    them.  Any mutation (the reference API allows callers to drop, merge or pad superpixels before pooling) forgets
    them, and `pool` then rebuilds the CSR from the dict's current contents."""
    assignment: Optional[SuperpixelAssignment] = None

    def _touch(self):
        self.assignment = None

    def __setitem__(self, k, v):
        self._touch()
        super().__setitem__(k, v)

    def __delitem__(self, k):
        self._touch()
        super().__delitem__(k)

    def pop(self, *a):
        self._touch()
        return super().pop(*a)

    def popitem(self):
        self._touch()
        return super().popitem()

    def update(self, *a, **k):
        self._touch()
        super().update(*a, **k)

    def setdefault(self, k, d=None):
        self._touch()
        return super().setdefault(k, d)

    def clear(self):
        self._touch()
        super().clear()

    def __ior__(self, other):
        self._touch()
        return super().__ior__(other)


class PatchToSuperpixelMapper:
    """Maps image patches to superpixels (reference: models/sppp.py:77-128)."""

    def __init__(self, patch_size: int = 16):
        self.patch_size = patch_size

    def assign_batch(self, segmentation_maps: torch.Tensor, img_size: int,
                     r_cap: Optional[int] = None) -> SuperpixelAssignment:
        """[B,H,W] integer label maps -> SuperpixelAssignment, one launch pair for the whole batch."""
        seg = segmentation_maps
        if seg.dtype != torch.int64:
            seg = seg.to(torch.int64)
        P = (img_size // self.patch_size) ** 2
        cap = P if r_cap is None else r_cap
        dom, slot, num_slots, counts, slot_label, offsets, order = ops.sppp_assign(seg, self.patch_size, img_size, cap)
        return SuperpixelAssignment(dom, slot, num_slots, counts, slot_label, offsets, order, cap)

    def assign_batch_with_centroids(self, segmentation_maps: torch.Tensor, img_size: int, r_cap: int, num_superpixels: int):
        """(SuperpixelAssignment, centroids fp32 [B,K,2]) from ONE pass over the label maps: `assign_batch` plus the
        superpixel centroids of models/sppp_mhla.py:226-262."""
        seg = segmentation_maps
        if seg.dtype != torch.int64:
            seg = seg.to(torch.int64)
        *a, cent = ops.sppp_assign_centroids(seg, self.patch_size, img_size, r_cap, num_superpixels)
        return SuperpixelAssignment(*a, r_cap), cent

    def map_patches(self, segmentation_map: torch.Tensor, img_size: int) -> Dict[int, List[int]]:
        """One image [H,W] -> {dominant label: [patch indices]} in first-seen order (sppp.py:91-128)."""
        a = self.assign_batch(segmentation_map.unsqueeze(0), img_size)
        d = a.to_dicts()[0]
        d.assignment = a
        d._sizes = [len(v) for v in d.values()]
        return d


def _assignment_from_dict(d: Dict[int, List[int]], num_patches: int, device) -> SuperpixelAssignment:
    """Host-side CSR for a dict that did not come from `map_patches` (slow path, API compatibility only)."""
    R = len(d)
    cap = max(R, 1)
    slot = torch.full((1, num_patches), -1, dtype=torch.int32)
    counts = torch.zeros((1, cap), dtype=torch.int32)
    offsets = torch.zeros((1, cap + 1), dtype=torch.int32)
    labels = torch.zeros((1, cap), dtype=torch.int64)
    flat: List[int] = []
    for r, (lab, patches) in enumerate(d.items()):
        labels[0, r] = int(lab)
        counts[0, r] = len(patches)
        offsets[0, r + 1] = offsets[0, r] + len(patches)
        for p in patches:
            if not 0 <= int(p) < num_patches:      # the reference's fancy indexing raises IndexError here (sppp.py:209)
                raise IndexError(f"index {int(p)} is out of bounds for dimension 0 with size {num_patches}")
            slot[0, p] = r
        flat.extend(int(p) for p in patches)
    offsets[0, R + 1:] = offsets[0, R]
    order = torch.zeros((1, max(num_patches, len(flat))), dtype=torch.int32)
    order[0, :len(flat)] = torch.tensor(flat, dtype=torch.int32)
    t = lambda x: x.to(device)
    return SuperpixelAssignment(t(torch.zeros((1, num_patches), dtype=torch.int64)), t(slot),
                                t(torch.tensor([R], dtype=torch.int32)), t(counts), t(labels), t(offsets), t(order), cap)


class SuperpixelPooling:
    """Pools patch embeddings based on superpixel regions (reference: models/sppp.py:131-223)."""

    def __init__(self, pooling_type: str = 'mean'):
        self.pooling_type = pooling_type

    def pool_batch(self, patch_embeddings: torch.Tensor, assignment: SuperpixelAssignment, num_superpixels: int,
                   validate: bool = True) -> torch.Tensor:
        """[B,P,D] -> fp32 [B,R,D] for the whole batch.  Like `torch.stack` at sppp_mhla.py:300 this needs every image
        to have exactly R = num_superpixels slots; with validate=True that is checked (one small D2H read) and a
        RuntimeError raised otherwise."""
        if self.pooling_type not in ('mean', 'max', 'attention'):
            raise ValueError(f"Unknown pooling type: {self.pooling_type}")
        if validate:
            ns = assignment.num_slots
            lo, hi = int(ns.min()), int(ns.max())
            if lo != hi or lo != num_superpixels:
                raise RuntimeError(f"stack expects each tensor to be equal size: images have between {lo} and {hi} "
                                   f"superpixel slots, expected {num_superpixels} (reference sppp_mhla.py:300)")
        a = assignment
        return self._pool_op(patch_embeddings, a, num_superpixels)

    def _pool_op(self, x: torch.Tensor, a: SuperpixelAssignment, R: int) -> torch.Tensor:
        if self.pooling_type == 'max':          # sppp.py:178-179 / 211-212
            return ops.sppp_pool_max(x, a.order, a.offsets, a.num_slots, R)[0]
        if self.pooling_type == 'attention':    # sppp.py:180-184 / 213-216
            return ops.sppp_pool_attn(x, a.order, a.offsets, a.num_slots, R)[0]
        return ops.sppp_pool(x, a.slot, a.counts, a.order, a.offsets, a.num_slots, R)

    def pool(self, patch_embeddings: torch.Tensor, superpixel_to_patches: Dict[int, List[int]]) -> torch.Tensor:
        """Reference signature: [N,D] (or [B,N,D] with one shared dict) + dict -> [R,D] (or [B,R,D]), fp32."""
        if self.pooling_type not in ('mean', 'max', 'attention'):
            raise ValueError(f"Unknown pooling type: {self.pooling_type}")
        x = patch_embeddings
        batched = x.dim() == 3
        xb = x if batched else x.unsqueeze(0)
        B, P, D = xb.shape
        a = getattr(superpixel_to_patches, "assignment", None)
        if a is not None:
            # the lists inside the dict can be edited in place without the dict noticing: the cached arrays are reused
            # only if the dict still has the sizes they were built from
            sizes = getattr(superpixel_to_patches, "_sizes", None)
            if sizes != [len(v) for v in superpixel_to_patches.values()]:
                a = None
        if a is None or a.slot.shape[1] != P or a.slot.device != x.device:
            a = _assignment_from_dict(superpixel_to_patches, P, x.device)
        R = len(superpixel_to_patches)
        if B > 1:  # one shared dict for the whole batch (3-D branch, sppp.py:159-191)
            ex = lambda t: t.expand(B, *t.shape[1:]).contiguous()
            a = SuperpixelAssignment(ex(a.dom), ex(a.slot), ex(a.num_slots), ex(a.counts), ex(a.slot_label),
                                     ex(a.offsets), ex(a.order), a.r_cap)
        if R == 0:
            return torch.zeros((B, 0, D) if batched else (0, D), device=x.device)
        out = self._pool_op(xb, a, R)
        return out if batched else out[0]
