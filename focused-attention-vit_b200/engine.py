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
    """cuda_graph=True captures the whole step (forward, loss, backward with the overlapped all-reduce, AdamW) into one
    CUDA graph after three eager warm-up steps (which also set up NCCL's communicator) and replays it afterwards: the launch-bound small workloads (SPPP +
    ViT-S: ~1300 launches of 5-10 us each) then run at GPU speed instead of Python speed.  Inputs are copied into the
    graph's static buffers; shapes must not change between calls."""

    def __init__(self, model: torch.nn.Module, lr: float = 1e-4, weight_decay: float = 0.05,
                 autocast_dtype: Optional[torch.dtype] = torch.bfloat16, process_group=None, bucket_mb: float = 32.0,
                 optimizer=True, cuda_graph: bool = False, dp_mode: str = "overlap", param_groups=None,
                 dp_grad_dtype: torch.dtype = torch.float32, backward_gemm_tiles: Optional[str] = None):
        """dp_mode (data parallel only): "overlap" = one coalesced all-reduce per gradient bucket, launched from the
        backward hooks and running beside the rest of backward (also inside the captured graph); "deferred" = one
        coalesced all-reduce of every gradient after backward (inside the graph when cuda_graph); "split" = deferred, with
        the collective issued eagerly between two graphs (backward | AdamW), for NCCL builds that cannot be captured."""
        if dp_mode not in ("overlap", "deferred", "split", "none"):
            raise ValueError(f"unknown dp_mode {dp_mode!r}")
        self.dp_mode = dp_mode
        self.model = model
        self.autocast_dtype = autocast_dtype
        # "none": DIAGNOSTIC ONLY (bench.py --dp-mode none): ranks train independently, no gradient exchange — what N ranks
        # cost without any collective (clock / power skew between the GPUs of a box)
        self.reducer = GradAllReducer(model.parameters(), bucket_mb=bucket_mb, process_group=process_group,
                                      enabled=False if dp_mode == "none" else None, grad_dtype=dp_grad_dtype)
        self.reducer.overlap = dp_mode == "overlap"
        # backward_gemm_tiles="steal": the backward GEMMs hand out their tiles by work stealing
        # (raw.gemm_tile_scheduler), so that a CTA pair which cannot be resident while the overlapped NCCL all-reduce
        # holds SMs does not run its whole share as a second wave.  Opt-in: at 8 GPUs it measured no better than the
        # static schedule (33.52 against 33.25 ms per step, profiles/r2_dp.md), and it costs ~1 % on a GPU of our own.
        if backward_gemm_tiles is None:
            backward_gemm_tiles = "static"
        if backward_gemm_tiles not in ("static", "steal"):
            raise ValueError(f"unknown backward_gemm_tiles {backward_gemm_tiles!r}")
        self.backward_gemm_tiles = backward_gemm_tiles
        # lr / weight decay: the reference's defaults (main.py:129-132).  optimizer: True / "favit" = the multi-tensor
        # AdamW kernel of this library (optim.FusedAdamW; `param_groups` = e.g. optim.reference_param_groups for the three
        # groups of experiments/mhla_pretrained.py:320-327), "torch" = torch.optim.AdamW(fused=True), False = none
        groups = param_groups if param_groups is not None else model.parameters()
        if optimizer in (True, "favit"):
            from .optim import FusedAdamW
            self.opt = FusedAdamW(groups, lr=lr, weight_decay=weight_decay)
        elif optimizer == "torch":
            self.opt = torch.optim.AdamW(groups, lr=lr, weight_decay=weight_decay, fused=True, capturable=cuda_graph)
        elif not optimizer:
            self.opt = None
        else:
            raise ValueError(f"unknown optimizer {optimizer!r}")
        self.cuda_graph = cuda_graph
        self._graph = None
        self._graph_opt = None
        self._static = None
        self._static_loss = None
        self._calls = 0

    def _fwd_bwd(self, images, labels, segmentation_maps):
        self.reducer.zero_grad()
        with torch.autocast("cuda", dtype=self.autocast_dtype or torch.bfloat16,
                            enabled=self.autocast_dtype is not None):
            logits = self.model(images) if segmentation_maps is None else self.model(images, segmentation_maps)
        loss = F.cross_entropy(logits.float(), labels)
        if self.backward_gemm_tiles == "steal":
            from . import raw
            raw.gemm_tile_scheduler("steal")        # read at launch (and baked into a captured graph)
            try:
                loss.backward()
            finally:
                raw.gemm_tile_scheduler("static")
        else:
            loss.backward()
        return loss.detach()

    def _eager(self, images, labels, segmentation_maps):
        loss = self._fwd_bwd(images, labels, segmentation_maps)
        self.reducer.finish()
        if self.opt is not None:
            self.opt.step()
        return loss

    def _capture(self, images, labels, segmentation_maps):
        """One graph for the whole step: zero_grad, forward, loss, backward, the gradient all-reduce(s) — forked onto
        NCCL's stream by the backward hooks and joined before the optimizer — and AdamW.  dp_mode "split" keeps the
        collective out of the capture: graph A = forward + backward, eager all-reduce, graph B = AdamW."""
        self._static = [images.clone(), labels.clone(), None if segmentation_maps is None else segmentation_maps.clone()]
        self._graph = torch.cuda.CUDAGraph()
        if not (self.reducer.enabled and self.dp_mode == "split"):
            # thread_local: NCCL's watchdog thread may query events while this thread captures
            with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
                self._static_loss = self._eager(*self._static)
            return
        with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
            self._static_loss = self._fwd_bwd(*self._static)
        if self.opt is not None:
            self._graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_opt, pool=self._graph.pool()):
                self.opt.step()

    def recapture(self) -> None:
        """Drop the captured graph(s); the next call captures the step again (bench.py re-captures with external event
        nodes around every favit launch to time the kernels inside replayed steps)."""
        self._graph = self._graph_opt = None
        self._static = self._static_loss = None
        self._calls = max(self._calls, 3)

    def __call__(self, images: torch.Tensor, labels: torch.Tensor,
                 segmentation_maps: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Device tensors in, detached scalar loss (device tensor) out."""
        if not self.cuda_graph:
            return self._eager(images, labels, segmentation_maps)
        self._calls += 1
        if self._calls <= 3:                      # eager warm-up: lazy initialisation must not be captured
            return self._eager(images, labels, segmentation_maps)
        if self._graph is None:
            torch.cuda.synchronize()
            self._capture(images, labels, segmentation_maps)
        for dst, src in zip(self._static, (images, labels, segmentation_maps)):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        if self.reducer.enabled and self.dp_mode == "split":
            self.reducer.finish()                 # one coalesced all-reduce of the gradients the graph left in place
            if self._graph_opt is not None:
                self._graph_opt.replay()
        return self._static_loss


def bind_host_to_gpu(index: int):
    """Pin the calling thread's CPU affinity to the cores NVML reports as local to CUDA device `index` (one process per
    GPU), so that pinned host buffers allocated afterwards land on the socket the GPU hangs off: a host -> device copy
    from the far socket crosses the inter-socket link and runs at a fraction of the PCIe rate.  Returns the previous
    affinity set (hand it to os.sched_setaffinity(0, ...) to undo), or None when NVML / the affinity call is unavailable
    — a performance hint only, nothing depends on it."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:      # CUDA_VISIBLE_DEVICES renumbers CUDA devices but not NVML's: go through the UUID
            handle = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(index).uuid))
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return before
    except Exception:
        return None


class HostBatchFeeder:
    """Pinned host batches -> device, one step ahead of the compute stream (double buffered), so that the H2D copy of
    step i+1 overlaps step i.  Every step's inputs still cross PCIe inside the timed region.  The two device slots are
    allocated once (no allocator traffic, no cross-stream frees in the steady state)."""

    def __init__(self, device, n_tensors: int):
        self.stream = torch.cuda.Stream(device=device)
        self.device = device
        self.slots = [None, None]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.i = 0

    def prefetch(self, host_tensors) -> None:
        k = self.i % 2
        if self.slots[k] is None:
            self.slots[k] = [None if t is None else torch.empty(t.shape, dtype=t.dtype, device=self.device)
                             for t in host_tensors]
            torch.cuda.current_stream().synchronize()     # first use only: the buffers exist before the side stream runs
        # the slot's previous reader (the step that consumed it) has completed: callers synchronise once per step
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for dst, src in zip(self.slots[k], host_tensors):
                if dst is not None:
                    dst.copy_(src, non_blocking=True)
            self.events[k].record(self.stream)
        self.i += 1

    def get(self, k: int):
        torch.cuda.current_stream().wait_event(self.events[k % 2])
        return self.slots[k % 2]
