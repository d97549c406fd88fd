"""ctypes binding of libfavit_b200.so — the C-ABI declared in include/favit.h.

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.  Nothing here (or
anywhere in this package) imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfavit_b200.so")

F32, BF16 = 0, 1
EPI_NONE, EPI_GELU, EPI_DGELU_MUL = 0, 1, 2

_lib = None

_vp, _i, _i64, _f, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64

_SIGS = {
    "favit_version": ([], _i),
    "favit_device_cc": ([], _i),
    "favit_last_error": ([], C.c_char_p),
    "favit_last_kernel": ([], C.c_char_p),
    "favit_launch_count": ([], _u64),
    "favit_mhla_attn_fwd": ([_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i64, _i64, _i64, _i, _f, _u64, _vp], _i),
    "favit_cast_bf16_batched": ([_i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), _vp], _i),
    "favit_copy_batched": ([_i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), _i, _i, _vp], _i),
    "favit_colsum": ([_vp, _i, _vp, _i, _i, _i64, _vp], _i),
    "favit_mhla_attn_bwd": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f,
                             _i64, _i64, _i64, _i, _f, _u64, _vp], _i),
    "favit_linear_fwd": ([_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _i, _i, _i, _i, _vp], _i),
    "favit_linear_dgrad": ([_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i, _i, _i, _vp], _i),
    "favit_linear_fwd_dropout": ([_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _i, _i, _i, _i, _f, _vp,
                                  _u64, _vp], _i),
    "favit_linear_dgrad_dropout": ([_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i, _i, _i, _f, _vp, _u64, _vp], _i),
    "favit_dropout_cast": ([_vp, _vp, _i, _vp, _i, _i, _f, _vp, _u64, _vp], _i),
    "favit_linear_wgrad": ([_vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i, _i, _vp], _i),
    "favit_gemm_bf16_raw": ([_vp, _i, _i64, _vp, _i, _i64, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp], _i),
    "favit_set_gemm_tile_scheduler": ([_i], _i),
    "favit_latent_fold_fwd": ([_vp] * 10 + [_i, _i, _i, _vp], _i),
    "favit_latent_fold_bwd": ([_vp] * 11 + [_i, _i, _vp], _i),
    "favit_latent_fold_fwd_batched": ([_i, C.POINTER(_vp), _i, _i, _i, _vp], _i),
    "favit_latent_fold_bwd_batched": ([_i, C.POINTER(_vp), _i, _i, _vp], _i),
    "favit_layernorm_fwd": ([_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _f, _vp], _i),
    "favit_layernorm_bwd": ([_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp], _i),
    "favit_sppp_assign": ([_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "favit_sppp_assign_centroids": ([_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp], _i),
    "favit_sppp_centroids": ([_vp, _i, _i, _i, _i, _vp, _vp, _vp], _i),
    "favit_sppp_pool_fwd": ([_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp], _i),
    "favit_sppp_pool_max_fwd": ([_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp], _i),
    "favit_sppp_pool_max_bwd": ([_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp], _i),
    "favit_sppp_pool_attn_fwd": ([_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp], _i),
    "favit_sppp_pool_attn_bwd": ([_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp], _i),
    "favit_patchify": ([_vp, _vp, _i, _i, _i, _i, _i, _vp], _i),
    "favit_sppp_embed_tokens": ([_vp, _vp, _vp, _vp, _i, _i, _i, _vp], _i),
    "favit_sppp_pool_pixels": ([_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp], _i),
    "favit_slic_grid": ([_i, _i, _i, C.POINTER(_i), C.POINTER(_i)], _i),
    "favit_slic_segment": ([_vp, _i, _i, _i, _i, _i, _f, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "favit_adamw_multi": ([_i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64),
                           C.POINTER(_f), C.POINTER(_f), _vp, C.c_double, C.c_double, _f, _f, _vp], _i),
    "favit_sppp_pool_bwd": ([_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp], _i),
}


def exported_symbols():
    """Names every build of the library must export (checked by the CPU test-suite against include/favit.h)."""
    return sorted(_SIGS)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python focused-attention-vit_b200/build.py` "
                "(there is no CPU or PyTorch fallback for the favit ops)")
        import torch  # noqa: F401  (loads libcudart.so.12, which the library links dynamically)
        l = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(l, name)       # AttributeError if the symbol is not exported
            fn.argtypes = args
            fn.restype = res
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().favit_last_error().decode(errors="replace")
        kind = {1: "bad argument", 2: "unsupported", 3: "CUDA error", 4: "workspace"}.get(rc, f"status {rc}")
        raise RuntimeError(f"{what}: {kind}: {msg}")


def last_kernel() -> str:
    """Name and variant of the kernel the last dispatching C call on this thread chose (tests assert on it)."""
    return lib().favit_last_kernel().decode(errors="replace")


def launch_count() -> int:
    return int(lib().favit_launch_count())


# ------------------------------------------------------------------------------------------------
# optional per-launch timing (bench.py's roofline): when `PROFILE` is a list, every C call is bracketed by CUDA
# events on the current torch stream and (family, role, algorithmic work, start, end) is appended to it.  While the
# stream is being captured the events are recorded as EXTERNAL event nodes of the graph, so that after every replay
# `start.elapsed_time(end)` is the duration of that launch inside the replayed step.
# `ROLE` names the layer a launch belongs to (set by fused_block: "qkv", "attn", "proj", "fc1", "fc2", "ln", "fold").
# ------------------------------------------------------------------------------------------------
PROFILE = None
ROLE = ""


def call(family: str, work: float, fn, *args) -> int:
    """Invoke a C-ABI function; `work` = algorithmic FLOPs (GEMM/attention) or bytes (SPPP) of this launch."""
    if PROFILE is None:
        return fn(*args)
    import torch
    ext = torch.cuda.is_current_stream_capturing()
    e0 = torch.cuda.Event(enable_timing=True, external=ext)
    e1 = torch.cuda.Event(enable_timing=True, external=ext)
    e0.record()
    rc = fn(*args)
    e1.record()
    PROFILE.append((family, ROLE, work, e0, e1))
    return rc
