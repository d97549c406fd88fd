"""Data-parallel gradient all-reduce overlapped with backward (one process per GPU, NCCL over NVLink 5 / NVSwitch).

The reference is single-process (no DDP); this is new work.  Parameters are grouped, in reverse
registration order (the order backward produces their gradients), into buckets of ~`bucket_mb`.  Gradients stay where
autograd puts them (`p.grad` is dropped to None every step, so AccumulateGrad adopts the tensor the backward kernel wrote:
no zero-fill, no `+=` pass, no gather copy); a post-accumulate-grad hook counts a bucket's gradients down and, when the
bucket is complete, launches ONE grouped all-reduce(AVG) over its tensors (ncclGroupStart .. ncclGroupEnd: a single NCCL
kernel for the whole bucket) on a side stream, ordered after the backward kernels that produced the bucket, while the
rest of backward keeps running on the compute stream.  `finish()` makes the compute stream wait for the outstanding
reductions before the optimizer reads the gradients.

On CUDA the collectives go through `NcclComm`: a communicator of our own (bootstrapped over the torch.distributed group,
which stays the plumbing: rendezvous, barriers, the bench's max-over-ranks), driven by direct `ncclAllReduce` calls on a
stream we own.  Everything is stream-ordered and no other thread touches the events, so the whole exchange can be
captured into the CUDA graph of the training step (engine.TrainStep): the collectives then sit in the graph as nodes
forked off the backward chain.  (ProcessGroupNCCL's watchdog / flight recorder query the events of captured collectives
from another thread, which is what makes capturing torch's own all-reduce fragile.)  With the gloo backend (CPU tests)
the same buckets go through the process group's `allreduce_coalesced`.

Every op on the path is per image (no BatchNorm, no cross-sample statistic), so averaged gradients of B/N-image shards
equal the gradient of the B-image batch with a mean loss.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

_NCCL_FLOAT32, _NCCL_BFLOAT16, _NCCL_AVG, _NCCL_SUM = 7, 9, 4, 0


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


class NcclComm:
    """A NCCL communicator over the ranks of a torch.distributed group (libnccl.so.2, the copy torch has loaded)."""

    def __init__(self, process_group=None):
        self.lib = C.CDLL("libnccl.so.2")
        self.lib.ncclGetErrorString.restype = C.c_char_p
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        self.lib.ncclAllReduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.ncclCommDestroy.argtypes = [C.c_void_p]
        self.rank = dist.get_rank(process_group)
        self.world = dist.get_world_size(process_group)
        uid = _UniqueId()
        if self.rank == 0:
            self._check(self.lib.ncclGetUniqueId(C.byref(uid)), "ncclGetUniqueId")
        box = [bytes(uid.internal) if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                                   group=process_group)
        C.memmove(C.byref(uid), box[0], 128)
        self.comm = C.c_void_p()
        self._check(self.lib.ncclCommInitRank(C.byref(self.comm), self.world, uid, self.rank), "ncclCommInitRank")
        self.stream = torch.cuda.Stream()

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise RuntimeError(f"{what} failed: {self.lib.ncclGetErrorString(rc).decode(errors='replace')}")

    def all_reduce_avg(self, tensors: List[torch.Tensor]) -> torch.cuda.Event:
        """In-place average of every tensor over the ranks, as ONE grouped launch on the communicator's stream, ordered
        after everything enqueued so far on the current stream.  Returns the event that marks its completion; the
        caller keeps the tensors alive until it has waited for it (they are the parameters' .grad)."""
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        st = C.c_void_p(self.stream.cuda_stream)
        self._check(self.lib.ncclGroupStart(), "ncclGroupStart")
        for t in tensors:
            if not t.is_contiguous() or t.dtype not in (torch.float32, torch.bfloat16):
                raise TypeError("NcclComm.all_reduce_avg takes contiguous fp32 / bf16 tensors")
            p = C.c_void_p(t.data_ptr())
            dt = _NCCL_FLOAT32 if t.dtype == torch.float32 else _NCCL_BFLOAT16
            self._check(self.lib.ncclAllReduce(p, p, t.numel(), dt, _NCCL_AVG, self.comm, st), "ncclAllReduce")
        self._check(self.lib.ncclGroupEnd(), "ncclGroupEnd")
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return ev

    def destroy(self) -> None:
        if self.comm:
            self.lib.ncclCommDestroy(self.comm)
            self.comm = C.c_void_p()


class GradAllReducer:
    SMALL = 65536      # gradients below this many elements are packed into one flat buffer per all-reduce

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 32.0, process_group=None,
                 enabled: Optional[bool] = None, grad_dtype: torch.dtype = torch.float32):
        """grad_dtype=torch.bfloat16 (NCCL only): the gradients cross NVLink as ONE flat bf16 buffer per all-reduce (half
        the bytes; packed from / unpacked into the fp32 .grad tensors by the multi-tensor copy kernel, the reduction
        itself accumulates in fp32 inside NCCL).  The averaged gradient is then rounded to bf16 — a data-parallel design
        choice like DDP's bf16 compression hook, off by default: fp32 keeps the 1e-5 equality with the large batch."""
        if grad_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("grad_dtype must be torch.float32 or torch.bfloat16")
        self.grad_dtype = grad_dtype
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
        self.overlap = True          # False: hooks are silent and finish() reduces everything in one grouped call
        self.collectives = 0         # all-reduce launches issued since construction
        self.nccl: Optional[NcclComm] = None
        self._flat = {}
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
        if dist.get_backend(process_group) == "nccl" and self.params and self.params[0].is_cuda:
            self.nccl = NcclComm(process_group)
        else:
            self._pg = process_group if process_group is not None else dist.distributed_c10d._get_default_group()
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

    def _reduce(self, tensors: List[torch.Tensor], key=None):
        """One grouped all-reduce.  On NCCL the SMALL tensors of the set (biases, LayerNorm parameters: ~100 of a ViT's
        ~150 gradients, a few KB each) travel as ONE flat buffer — packed and unpacked by a multi-tensor copy kernel —
        because every member of a NCCL group is a collective of its own with its own latency; the large weight
        gradients are reduced in place.  Returns (completion handle, tensors, unpack job or None)."""
        self.collectives += 1
        if self.nccl is not None:
            small = tensors if self.grad_dtype == torch.bfloat16 else [t for t in tensors if t.numel() < self.SMALL]
            if len(small) < 2 and self.grad_dtype == torch.float32:
                return self.nccl.all_reduce_avg(tensors), tensors, None
            from . import raw
            flat = self._flat.get(key)
            total = sum((t.numel() + 7) // 8 * 8 for t in small)          # 16-byte aligned views
            if flat is None or flat.numel() != total or flat.dtype != self.grad_dtype:
                flat = torch.empty(total, dtype=self.grad_dtype, device=small[0].device)
                self._flat[key] = flat
            views, off = [], 0
            for t in small:
                views.append(flat[off:off + t.numel()])
                off += (t.numel() + 7) // 8 * 8
            raw.copy_batched(small, views)
            big = [] if self.grad_dtype == torch.bfloat16 else [t for t in tensors if t.numel() >= self.SMALL]
            return self.nccl.all_reduce_avg(big + [flat]), tensors, (views, small)
        opts = dist.AllreduceCoalescedOptions()
        opts.reduceOp = dist.ReduceOp.SUM          # gloo has no AVG: sum and scale afterwards
        return self._pg.allreduce_coalesced(tensors, opts), tensors, None

    def _hook(self, p: torch.nn.Parameter) -> None:
        if not self.overlap:
            return
        bi = self._bucket_of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._works.append(self._reduce([q.grad for q in self.buckets[bi]], key=bi))

    def finish(self) -> None:
        """Reduce what the hooks have not (deferred mode: everything, in one grouped call; overlapped mode: buckets
        with parameters that got no gradient this step), then wait — stream-wise on CUDA — for every outstanding
        reduction.  Parameters without a gradient are skipped; they must be the same on every rank."""
        if not self.enabled:
            return
        if not self.overlap:
            left = [p.grad for p in self.params if p.grad is not None]
            if left:
                self._works.append(self._reduce(left, key="all"))
        else:
            for bi, n in enumerate(self._pending):
                if n != 0:
                    left = [p.grad for p in self.buckets[bi] if p.grad is not None]
                    if left:
                        self._works.append(self._reduce(left, key=("rest", bi)))
        self._pending = [0] * len(self._pending)
        for w, tensors, unpack in self._works:
            if self.nccl is not None:
                torch.cuda.current_stream().wait_event(w)
                if unpack is not None:
                    from . import raw
                    raw.copy_batched(*unpack)           # flat buffer -> the small gradients, on the compute stream
            else:
                w.wait()
                torch._foreach_div_(tensors, float(self.world))
        self._works = []

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
        if self.nccl is not None:
            self.nccl.destroy()
            self.nccl = None

    @property
    def grad_bytes(self) -> int:
        return sum(p.numel() * 4 for p in self.params)
