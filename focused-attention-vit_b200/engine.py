"""Training / inference step around the drop-in models: the call a user makes (bench.py's `e2e` goes through it).

One step = zero gradients -> forward (bf16 autocast by default) -> cross-entropy -> backward, with the bucketed
gradient all-reduce of `dp.GradAllReducer` overlapped with backward when a process group is up -> AdamW.
Mirrors the reference's hot loop (experiments/mhla_pretrained.py:359-369: H2D copy, zero_grad, forward, loss,
backward, step, loss.item()) with the data-parallel all-reduce the reference does not have.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from .dp import GradAllReducer


class TrainStep:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-4, weight_decay: float = 0.05,
                 autocast_dtype: Optional[torch.dtype] = torch.bfloat16, process_group=None, bucket_mb: float = 32.0,
                 optimizer: bool = True):
        self.model = model
        self.autocast_dtype = autocast_dtype
        self.reducer = GradAllReducer(model.parameters(), bucket_mb=bucket_mb, process_group=process_group)
        # lr / weight decay: the reference's defaults (main.py:129-132)
        self.opt = (torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=weight_decay, fused=True)
                    if optimizer else None)

    def __call__(self, images: torch.Tensor, labels: torch.Tensor,
                 segmentation_maps: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Device tensors in, detached scalar loss (device tensor) out."""
        self.reducer.zero_grad()
        with torch.autocast("cuda", dtype=self.autocast_dtype or torch.bfloat16,
                            enabled=self.autocast_dtype is not None):
            logits = self.model(images) if segmentation_maps is None else self.model(images, segmentation_maps)
        loss = F.cross_entropy(logits.float(), labels)
        loss.backward()
        self.reducer.finish()
        if self.opt is not None:
            self.opt.step()
        return loss.detach()


class HostBatchFeeder:
    """Pinned host batches -> device, one step ahead of the compute stream (double buffered), so that the H2D copy of
    step i+1 overlaps step i.  Every step's inputs still cross PCIe inside the timed region."""

    def __init__(self, device, n_tensors: int):
        self.stream = torch.cuda.Stream(device=device)
        self.device = device
        self.slots = [None, None]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.i = 0

    def prefetch(self, host_tensors) -> None:
        k = self.i % 2
        with torch.cuda.stream(self.stream):
            self.slots[k] = [None if t is None else t.to(self.device, non_blocking=True) for t in host_tensors]
            self.events[k].record(self.stream)
        self.i += 1

    def get(self, k: int):
        torch.cuda.current_stream().wait_event(self.events[k % 2])
        out = self.slots[k % 2]
        for t in out:
            if t is not None:
                t.record_stream(torch.cuda.current_stream())
        return out
