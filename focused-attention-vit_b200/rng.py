"""Device-resident seed for the counter-based dropout masks of the fused kernels.

A fused forward call takes `call_seed(device)`: a private int64 [1] copy of the per-device state (saved for its
backward, which regenerates the same masks) and bumps the state IN STREAM.  Because both are stream-ordered device
operations, a CUDA graph that captures the training step draws a fresh mask on every replay — a host integer baked
into kernel arguments would replay the same mask forever.  The state starts from torch's CPU generator, so
`torch.manual_seed` makes runs reproducible.
"""
from __future__ import annotations

import torch

_STATE = {}
_GOLDEN = 0x9E3779B97F4A7C15 - (1 << 64)      # as a signed 64-bit increment


def _state(device: torch.device) -> torch.Tensor:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    st = _STATE.get(key)
    if st is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("favit_b200.rng: the dropout seed state must exist before CUDA-graph capture "
                               "(run one eager step first, as engine.TrainStep does)")
        st = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).to(device)
        _STATE[key] = st
    return st


def call_seed(device: torch.device) -> torch.Tensor:
    st = _state(device)
    seed = st.clone()
    st.add_(_GOLDEN)
    return seed


def set_state(device: torch.device, value: int) -> None:
    """Tests: pin the state so that the oracle can regenerate the masks of the next forward call."""
    _state(device).fill_(int(value))
