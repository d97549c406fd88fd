"""`FusedAdamW`: torch.optim.AdamW semantics on favit's multi-tensor kernel (csrc/adamw.cu).

Replaces the optimizer.step() of the reference's training loop (experiments/mhla_pretrained.py:320-327, 367: three
parameter groups, `latent_proj` at five times the base learning rate; main.py:129-132).  Every parameter group keeps
its own lr / weight_decay (betas / eps may differ per group too: one launch set per distinct pair); the step count lives
on the device and is advanced in-stream, so the optimizer can be captured in a CUDA graph; `grad_scale` folds the 1/world
of a summing gradient all-reduce into the update."""
from __future__ import annotations

import ctypes as C
from typing import Iterable

import torch

from . import _lib as L


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, grad_scale: float = 1.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = grad_scale
        self._step_t = None

    def _state(self, p: torch.Tensor):
        st = self.state[p]
        if not st:
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def init_state(self) -> None:
        """Allocate moments and the device step counter up front (before CUDA-graph capture)."""
        for group in self.param_groups:
            for p in group["params"]:
                if p.requires_grad:
                    self._state(p)
                    if self._step_t is None:
                        self._step_t = torch.zeros(1, dtype=torch.int64, device=p.device)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._step_t is None:
            self.init_state()
        if self._step_t is None:
            return loss
        self._step_t.add_(1)
        by_hp = {}
        for group in self.param_groups:
            key = (float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]))
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise TypeError("FusedAdamW updates fp32 CUDA parameters with fp32 gradients (sm_100a kernel, no fallback)")
                if p.grad.is_sparse or not p.is_contiguous():
                    raise RuntimeError("FusedAdamW needs dense, contiguous parameters")
                st = self._state(p)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                by_hp.setdefault(key, []).append((p, g, st["exp_avg"], st["exp_avg_sq"], float(group["lr"]),
                                                  float(group["weight_decay"])))
        stream = torch.cuda.current_stream().cuda_stream
        for (b1, b2, eps), items in by_hp.items():
            n = len(items)
            arr = lambda k: (C.c_void_p * n)(*[it[k].data_ptr() for it in items])
            rc = L.call("adamw", float(sum(it[0].numel() for it in items)) * 28.0, L.lib().favit_adamw_multi, n, arr(0), arr(1),
                        arr(2), arr(3), (C.c_int64 * n)(*[it[0].numel() for it in items]),
                        (C.c_float * n)(*[it[4] for it in items]), (C.c_float * n)(*[it[5] for it in items]),
                        self._step_t.data_ptr(), b1, b2, eps, float(self.grad_scale), stream)
            L.check(rc, "favit_adamw_multi")
        return loss


def reference_param_groups(model: torch.nn.Module, lr: float, head_lr: float, latent_lr_mult: float = 5.0):
    """The three AdamW parameter groups of the reference's training loop (experiments/mhla_pretrained.py:320-327):
    everything but the head and the latent projections at `lr`, `latent_proj` at 5 x lr, the head at `head_lr`."""
    named = list(model.named_parameters())
    return [
        {"params": [p for n, p in named if "head" not in n and "latent_proj" not in n], "lr": lr},
        {"params": [p for n, p in named if "latent_proj" in n], "lr": lr * latent_lr_mult},
        {"params": list(model.head.parameters()), "lr": head_lr},
    ]
