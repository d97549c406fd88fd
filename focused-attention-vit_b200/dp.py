"""Data-parallel gradient all-reduce overlapped with backward (one process per GPU, NCCL over NVLink 5 / NVSwitch).

The reference is single-process (no DDP); this is new work.  Parameters are grouped, in reverse
registration order (the order backward produces their gradients), into buckets of ~`bucket_mb`.  Gradients stay where
autograd puts them (`p.grad` is dropped to None every step, so AccumulateGrad adopts the tensor the backward kernel wrote:
no zero-fill, no `+=` pass, no gather copy); a post-accumulate-grad hook counts a bucket's gradients down and, when the
bucket is complete, launches ONE coalesced all-reduce(AVG) over its tensors (a single NCCL group, i.e. one NCCL kernel for
the whole bucket) on NCCL's own stream, ordered after the backward kernels that produced the bucket, while the rest of
backward keeps running on the compute stream.  `finish()` makes the compute stream wait for the outstanding reductions
before the optimizer reads the gradients.  All of it is stream-ordered, so it can be captured into the CUDA graph of the
training step (engine.TrainStep): the collectives then sit in the graph as nodes forked off the backward chain.

Every op on the path is per image (no BatchNorm, no cross-sample statistic), so averaged gradients of B/N-image shards
equal the gradient of the B-image batch with a mean loss.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 32.0, process_group=None,
                 enabled: Optional[bool] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = process_group
        if enabled is None:
            enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self.enabled = enabled
        self.world = dist.get_world_size(process_group) if self.enabled else 1
        self.buckets: List[List[torch.nn.Parameter]] = []
        self._bucket_of = {}
        self._pending: List[int] = []
        self._works = []
        self._handles = []
        self.overlap = True          # False: hooks are silent and finish() reduces everything in one coalesced call
        self.collectives = 0         # all-reduce calls issued since construction (bench.py reports it per step)
        if not self.enabled:
            return
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError("GradAllReducer expects fp32 master parameters")
        cap = int(bucket_mb * (1 << 20)) // 4
        cur: List[torch.nn.Parameter] = []
        n = 0
        for p in reversed(self.params):
            if cur and n + p.numel() > cap:
                self.buckets.append(cur)
                cur, n = [], 0
            cur.append(p)
            n += p.numel()
        if cur:
            self.buckets.append(cur)
        for bi, b in enumerate(self.buckets):
            for p in b:
                self._bucket_of[p] = bi
        self._pending = [len(b) for b in self.buckets]
        self._pg = process_group if process_group is not None else dist.distributed_c10d._get_default_group()
        # gloo has no AVG: sum and scale afterwards
        self._avg = dist.get_backend(process_group) == "nccl"
        for p in self.params:
            self._handles.append(p.register_post_accumulate_grad_hook(self._hook))

    # -- per step ---------------------------------------------------------------------------------
    def zero_grad(self) -> None:
        """Gradients are dropped, not zeroed: autograd then hands its freshly produced tensors over instead of adding
        them into zero-filled buffers (single process and data parallel alike)."""
        for p in self.params:
            p.grad = None
        if self.enabled:
            self._pending = [len(b) for b in self.buckets]
            self._works = []

    def _reduce(self, tensors: List[torch.Tensor]):
        opts = dist.AllreduceCoalescedOptions()
        opts.reduceOp = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self.collectives += 1
        return self._pg.allreduce_coalesced(tensors, opts), tensors

    def _hook(self, p: torch.nn.Parameter) -> None:
        if not self.overlap:
            return
        bi = self._bucket_of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._works.append(self._reduce([q.grad for q in self.buckets[bi]]))

    def finish(self) -> None:
        """Reduce what the hooks have not (deferred mode: everything, in one coalesced call; overlapped mode: buckets
        with parameters that got no gradient this step), then wait — stream-wise on CUDA — for every outstanding
        reduction.  Parameters without a gradient are skipped; they must be the same on every rank."""
        if not self.enabled:
            return
        if not self.overlap:
            left = [p.grad for p in self.params if p.grad is not None]
            if left:
                self._works.append(self._reduce(left))
        else:
            for bi, n in enumerate(self._pending):
                if n != 0:
                    left = [p.grad for p in self.buckets[bi] if p.grad is not None]
                    if left:
                        self._works.append(self._reduce(left))
        self._pending = [0] * len(self._pending)
        for w, tensors in self._works:
            w.wait()
            if not self._avg:
                torch._foreach_div_(tensors, float(self.world))
        self._works = []

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []

    @property
    def grad_bytes(self) -> int:
        return sum(p.numel() * 4 for p in self.params)
