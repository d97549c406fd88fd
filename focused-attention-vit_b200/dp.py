"""Data-parallel gradient all-reduce overlapped with backward (one process per GPU, NCCL over NVLink 5 / NVSwitch).

The reference is single-process (no DDP); this is new work.  Parameters are packed, in reverse
registration order (the order backward produces their gradients), into flat fp32 buckets; every `param.grad` is a view
into its bucket, so there is no gather copy.  A post-accumulate-grad hook counts a bucket's gradients down and, when the
bucket is complete, launches an asynchronous all-reduce(AVG) on it: NCCL runs on its own stream, ordered after the
backward kernels that produced the bucket, while the rest of backward keeps running on the compute stream.
`finish()` makes the compute stream wait for the outstanding reductions before the optimizer reads the gradients.

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
        self.buckets: List[torch.Tensor] = []
        self._bucket_of = {}
        self._size = []
        self._pending = []
        self._works = []
        self._handles = []
        self.overlap = True
        if not self.enabled:
            # single process: no flat buckets; gradients are dropped (set to None) every step so that autograd hands
            # its freshly produced tensors over instead of adding them into zero-filled buffers
            return
        cap = int(bucket_mb * (1 << 20)) // 4
        self._bucket_of = {}
        self._size = []
        cur: List[torch.nn.Parameter] = []
        n = 0
        groups = []
        for p in reversed(self.params):
            if cur and n + p.numel() > cap:
                groups.append(cur)
                cur, n = [], 0
            cur.append(p)
            n += p.numel()
        if cur:
            groups.append(cur)
        # every bucket is a slice of ONE flat buffer: the overlapped mode reduces bucket by bucket as backward fills
        # them, the deferred mode (backward replayed from a CUDA graph) reduces the whole buffer in a single call
        # instead of paying the launch latency of a dozen collectives after the step
        sizes = [sum(p.numel() for p in g) for g in groups]
        starts = [0]
        for sz in sizes:
            starts.append(starts[-1] + (sz + 63) // 64 * 64)       # 256-byte aligned bucket starts
        self.flat_all = torch.zeros(starts[-1], dtype=torch.float32, device=groups[0][0].device)
        for bi, g in enumerate(groups):
            flat = self.flat_all[starts[bi]:starts[bi] + sizes[bi]]
            off = 0
            for p in g:
                if p.dtype != torch.float32:
                    raise TypeError("GradAllReducer expects fp32 master parameters")
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                self._bucket_of[p] = bi
            self.buckets.append(flat)
            self._size.append(len(g))
        self._pending = list(self._size)
        self._works = []
        self._handles = []
        self.overlap = True
        if self.enabled:
            # gloo has no AVG: sum and scale afterwards
            self._avg = dist.ReduceOp.AVG if dist.get_backend(process_group) == "nccl" else None
            for p in self.params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook))

    # -- per step ---------------------------------------------------------------------------------
    def zero_grad(self) -> None:
        if not self.enabled:
            for p in self.params:
                p.grad = None
            return
        self.flat_all.zero_()
        self._pending = list(self._size)
        self._works = []

    def _hook(self, p: torch.nn.Parameter) -> None:
        if not self.overlap:       # deferred mode (backward replayed from a CUDA graph): finish() reduces every bucket
            return
        bi = self._bucket_of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            flat = self.buckets[bi]
            if self._avg is not None:
                w = dist.all_reduce(flat, op=self._avg, group=self.group, async_op=True)
            else:
                w = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._works.append((w, bi))

    def finish(self) -> None:
        """Wait (stream-wise on CUDA) for every outstanding bucket; reduce buckets whose hooks did not all fire
        (parameters unused in this step keep a zero gradient)."""
        if not self.enabled:
            return
        op = self._avg if self._avg is not None else dist.ReduceOp.SUM
        if not self.overlap:
            dist.all_reduce(self.flat_all, op=op, group=self.group, async_op=True).wait()
            if self._avg is None:
                self.flat_all.div_(self.world)
            self._pending = [0] * len(self._pending)
            self._works = []
            return
        for bi, left in enumerate(self._pending):
            if left != 0:
                op = self._avg if self._avg is not None else dist.ReduceOp.SUM
                self._works.append((dist.all_reduce(self.buckets[bi], op=op, group=self.group, async_op=True), bi))
                self._pending[bi] = 0
        for w, bi in self._works:
            w.wait()
            if self._avg is None:
                self.buckets[bi].div_(self.world)
        self._works = []

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []

    @property
    def grad_bytes(self) -> int:
        return sum(p.numel() * 4 for p in self.params)
