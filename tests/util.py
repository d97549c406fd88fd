"""Shared helpers for the test-suite."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Parity bars of BASELINE.json:north_star — about 1e-4 relative in fp32, about 2e-2 in bf16.
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def rel_err(got: torch.Tensor, ref: torch.Tensor, floor: float = 1e-30) -> float:
    """max |got - ref| / max(max |ref|, floor)  (fp64, on the CPU).  `floor` is the magnitude below which the reference
    counts as zero (e.g. a gradient that is analytically 0: the scale of its sibling gradients)."""
    g = got.detach().double().cpu()
    r = ref.detach().double().cpu()
    assert g.shape == r.shape, (g.shape, r.shape)
    denom = max(r.abs().max().item(), floor)
    return (g - r).abs().max().item() / denom


def assert_close(got, ref, dtype, what="", factor=1.0, floor=1e-30):
    e = rel_err(got, ref, floor)
    assert e <= TOL[dtype] * factor, f"{what}: rel err {e:.3e} > {TOL[dtype] * factor:.1e} ({dtype})"
