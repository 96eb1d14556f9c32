"""Precision modes and quantisers — API of the reference's quantization.py, executed by sm_100a kernels.

Same names, arguments, defaults and return types as `/root/reference/quantization.py`
(PrecisionMode :10-18, quantize_distance_squared :21-71, _grid_quantize :74-88,
_grid_quantize_safe :91-127, quantize_force :130-157, get_mode_from_string :160-175,
describe_mode :178-189).  The tensor functions take CUDA float32/float64 tensors of any shape and
run two kernels each (global min/max reduction, then the snap); the reference's host-side
`if max - min < 1e-10` branches become device-side flags, so no call synchronises with the host.
Inside `GalaxySimulation` these semantics are fused into the force kernel instead (see simulation.py).
"""
from __future__ import annotations

from enum import Enum
from typing import Optional

import torch

from . import _lib as L


class PrecisionMode(Enum):
    """quantization.py:10-18 — identical member names and values."""
    FLOAT64 = "float64"
    FLOAT32 = "float32"
    BFLOAT16 = "bfloat16"
    FLOAT16 = "float16"
    INT8_SIM = "int8_sim"
    INT4_SIM = "int4_sim"
    CUSTOM = "custom"


_GRID_LEVELS = {PrecisionMode.INT8_SIM: 256, PrecisionMode.INT4_SIM: 16}


def levels_for_mode(mode: PrecisionMode, custom_levels: Optional[int] = None) -> Optional[int]:
    """Grid levels of the log-space d² quantiser for `mode` (quantization.py:58-68); None for float modes."""
    if mode in _GRID_LEVELS:
        return _GRID_LEVELS[mode]
    if mode == PrecisionMode.CUSTOM:
        return custom_levels or 64
    return None


def _prepared(t: torch.Tensor) -> torch.Tensor:
    L.require_cuda(t)
    L.dtype_code(t)                      # raises for anything but fp32 / fp64
    return t.contiguous()


def _new_scalars(device) -> torch.Tensor:
    s = torch.empty(L.SCALAR_SLOTS, dtype=torch.int64, device=device)
    L.check(L.load().nb_reset_scalars(L.ptr(s), L.stream_ptr(device)), "nb_reset_scalars")
    return s


def _grid_quantize(tensor: torch.Tensor, levels: int) -> torch.Tensor:
    """Linear grid between the global min and max — quantization.py:74-88."""
    src = _prepared(tensor)
    if src.numel() == 0:
        return tensor
    lib, st = L.load(), L.stream_ptr(src.device)
    with torch.cuda.device(src.device):
        scal = _new_scalars(src.device)
        out = torch.empty_like(src)
        code = L.dtype_code(src)
        L.check(lib.nb_tensor_minmax(L.ptr(src), src.numel(), code, 0, 0.0, L.ptr(scal), st), "nb_tensor_minmax")
        L.check(lib.nb_grid_quantize(L.ptr(src), L.ptr(out), src.numel(), code, int(levels), L.ptr(scal), st),
                "nb_grid_quantize")
    return out.view(tensor.shape)


def _grid_quantize_safe(tensor: torch.Tensor, levels: int, min_val: float = 0.01,
                        return_index: bool = False):
    """Log-space grid above a floor — quantization.py:91-127.

    `return_index=True` (extension used by the parity tests) also returns the int32 level index
    round(normalized) of every element.
    """
    src = _prepared(tensor)
    if src.numel() == 0:
        return (tensor, None) if return_index else tensor
    lib, st = L.load(), L.stream_ptr(src.device)
    with torch.cuda.device(src.device):
        scal = _new_scalars(src.device)
        out = torch.empty_like(src)
        idx = torch.empty(src.shape, dtype=torch.int32, device=src.device) if return_index else None
        code = L.dtype_code(src)
        L.check(lib.nb_tensor_minmax(L.ptr(src), src.numel(), code, 1, float(min_val), L.ptr(scal), st),
                "nb_tensor_minmax")
        L.check(lib.nb_grid_quantize_safe(L.ptr(src), L.ptr(out), L.ptr(idx), src.numel(), code, int(levels),
                                          float(min_val), L.ptr(scal), st), "nb_grid_quantize_safe")
    out = out.view(tensor.shape)
    return (out, idx.view(tensor.shape)) if return_index else out


def _round_trip(tensor: torch.Tensor, mode: PrecisionMode) -> torch.Tensor:
    """`.half().float()` / `.bfloat16().float()` — quantization.py:53,56,146,149 (result is float32)."""
    src = _prepared(tensor)
    out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    if src.numel():
        with torch.cuda.device(src.device):
            L.check(L.load().nb_round_trip(L.ptr(src), L.ptr(out), src.numel(), L.dtype_code(src),
                                           L.MODE_CODES[mode.value], L.stream_ptr(src.device)), "nb_round_trip")
    return out.view(tensor.shape)


def snap_index(normalized: torch.Tensor) -> torch.Tensor:
    """round-half-to-even of already-normalised grid coordinates (the snap of quantization.py:85,120)."""
    src = _prepared(normalized)
    out = torch.empty(src.shape, dtype=torch.int32, device=src.device)
    if src.numel():
        with torch.cuda.device(src.device):
            L.check(L.load().nb_snap_index(L.ptr(src), L.ptr(out), src.numel(), L.dtype_code(src),
                                           L.stream_ptr(src.device)), "nb_snap_index")
    return out.view(normalized.shape)


def quantize_distance_squared(dist_sq: torch.Tensor, mode: PrecisionMode, custom_levels: int = None,
                              min_dist_sq: float = 0.01) -> torch.Tensor:
    """Precision degradation of d² — quantization.py:21-71."""
    if mode == PrecisionMode.FLOAT64:
        L.require_cuda(dist_sq)
        return dist_sq.double()                      # exact widening (dtype plumbing)
    if mode == PrecisionMode.FLOAT32:
        L.require_cuda(dist_sq)
        return dist_sq.float()
    if mode in (PrecisionMode.BFLOAT16, PrecisionMode.FLOAT16):
        return _round_trip(dist_sq, mode)
    levels = levels_for_mode(mode, custom_levels)
    if levels is not None:
        return _grid_quantize_safe(dist_sq, levels=levels, min_val=min_dist_sq)
    return dist_sq


def quantize_force(force: torch.Tensor, mode: PrecisionMode, custom_levels: int = None) -> torch.Tensor:
    """Optional quantisation of the force values — quantization.py:130-157."""
    if mode in (PrecisionMode.FLOAT64, PrecisionMode.FLOAT32):
        return force
    if mode in (PrecisionMode.BFLOAT16, PrecisionMode.FLOAT16):
        return _round_trip(force, mode)
    if mode == PrecisionMode.INT8_SIM:
        return _grid_quantize(force, levels=256)
    if mode == PrecisionMode.INT4_SIM:
        return _grid_quantize(force, levels=16)
    if mode == PrecisionMode.CUSTOM:
        return _grid_quantize(force, levels=custom_levels or 64)
    return force


_MODE_ALIASES = {
    "float64": PrecisionMode.FLOAT64, "float32": PrecisionMode.FLOAT32,
    "bfloat16": PrecisionMode.BFLOAT16, "bf16": PrecisionMode.BFLOAT16,
    "float16": PrecisionMode.FLOAT16, "fp16": PrecisionMode.FLOAT16,
    "int8": PrecisionMode.INT8_SIM, "int8_sim": PrecisionMode.INT8_SIM,
    "int4": PrecisionMode.INT4_SIM, "int4_sim": PrecisionMode.INT4_SIM,
    "custom": PrecisionMode.CUSTOM,
}


def get_mode_from_string(mode_str: str) -> PrecisionMode:
    """Case-insensitive lookup; unknown strings silently mean FLOAT64 — quantization.py:160-175."""
    return _MODE_ALIASES.get(mode_str.lower(), PrecisionMode.FLOAT64)


_DESCRIPTIONS = {
    PrecisionMode.FLOAT64: "64-bit float (baseline)",
    PrecisionMode.FLOAT32: "32-bit float (standard GPU)",
    PrecisionMode.BFLOAT16: "Brain Float 16 (AI precision, fast on RTX)",
    PrecisionMode.FLOAT16: "16-bit float (half precision)",
    PrecisionMode.INT8_SIM: "Simulated 8-bit (256 levels)",
    PrecisionMode.INT4_SIM: "Simulated 4-bit (16 levels)",
    PrecisionMode.CUSTOM: "Custom quantization levels",
}


def describe_mode(mode: PrecisionMode) -> str:
    """Human-readable description — quantization.py:178-189 (same strings)."""
    return _DESCRIPTIONS.get(mode, "Unknown mode")
