"""favit_b200 — B200-native (sm_100a) MHLA + SPPP hot path of zser092/Focused-Attention-ViT.

Import name: `favit_b200` (see the shim `favit_b200.py` at the repo root).  The package is a thin PyTorch host
over `libfavit_b200.so`; it has no CPU fallback and never imports `oracle/`.
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401

__all__ = ["_lib", "ops"]
