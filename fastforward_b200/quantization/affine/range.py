"""Integer bounds and range <-> (scale, offset) conversions
(reference: quantization/affine/range.py:9-122)."""

from __future__ import annotations

from typing import Optional, Tuple

import torch

from ... import ops


def integer_minimum(num_bits: float) -> float:
    return -(2 ** (num_bits - 1))


def integer_maximum(num_bits: float) -> float:
    return -integer_minimum(num_bits) - 1


def quantization_range(scale, offset, num_bits: float):
    """``((int_min + offset) * scale, (int_max + offset) * scale)`` -- tiny parameter-sized torch
    expressions, evaluated wherever the parameters live (range.py:31-51)."""
    offset = 0.0 if offset is None else offset
    return (integer_minimum(num_bits) + offset) * scale, (integer_maximum(num_bits) + offset) * scale


def _as_tensor(v, device) -> torch.Tensor:
    return v if isinstance(v, torch.Tensor) else torch.tensor(v, device=device)


def parameters_for_range(
    min_range, max_range, num_bits: float, symmetric: bool, allow_one_sided: bool, *, exact_none: bool = False
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """(scale, offset) that best represent [min_range, max_range], computed on the device in fp32.

    The reference decides "one-sided" with ``min_range.min() >= 0`` on the host (range.py:100, a
    device sync per call).  Here the decision is taken inside the kernel, so for a symmetric
    quantizer that allows one-sided ranges the returned offset is a tensor in both cases: zeros
    when the two-sided branch was taken (the value the reference stores in the offset buffer,
    nn/linear_quantizer.py:353-357) instead of ``None``.  ``exact_none=True`` restores the
    reference's ``None`` at the price of one host sync.
    """
    device = min_range.device if isinstance(min_range, torch.Tensor) else (
        max_range.device if isinstance(max_range, torch.Tensor) else torch.device("cuda"))
    mn, mx = _as_tensor(min_range, device), _as_tensor(max_range, device)
    shape = torch.broadcast_shapes(mn.shape, mx.shape)
    mn, mx = mn.expand(shape), mx.expand(shape)
    scale = torch.empty(shape, dtype=torch.float32, device=mn.device)
    has_offset = not (symmetric and not allow_one_sided)
    offset = torch.empty(shape, dtype=torch.float32, device=mn.device) if has_offset else None
    if scale.numel():
        ops.parameters_for_range_(mn, mx, num_bits, symmetric, allow_one_sided, scale, offset)
    if exact_none and symmetric and allow_one_sided and not bool(mn.min() >= 0):
        offset = None
    return scale, offset
