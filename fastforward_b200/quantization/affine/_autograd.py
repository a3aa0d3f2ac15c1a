"""Autograd bindings of the hot-path ops (reference: quantization/affine/_autograd.py:27-156).

Design rule kept from the reference: every gradient lives in *quantize*'s backward
(straight-through estimator with per-tile scale/offset sums); dequantize's backward is the
identity.  ``FakeQuantizeAffine`` is the fused quantize->dequantize pair (one HBM pass forward,
one backward) used where the reference runs the two back to back (export mode, weight fusing)."""

from __future__ import annotations

from typing import Any, Optional

import torch

from ... import ops


def _tensor_or_none(v, dtype: torch.dtype, device: torch.device):
    if v is None or isinstance(v, torch.Tensor):
        return v
    return torch.tensor(v, dtype=dtype, device=device)


def _float_dtype(data: torch.Tensor) -> torch.dtype:
    return data.dtype if data.dtype.is_floating_point else torch.get_default_dtype()


def _resolve_tile(data: torch.Tensor, tile_size):
    return tuple(data.shape) if isinstance(tile_size, str) else tuple(tile_size)


class QuantizeStaticAffine(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, data, scale, offset, tile_size, num_bits, quantized_dtype):
        ctx.save_for_backward(data, scale, offset)
        ctx.tile_size = _resolve_tile(data, tile_size)
        ctx.num_bits = num_bits
        return ops.quantize_by_tile(data, scale, ctx.tile_size, num_bits, quantized_dtype or data.dtype, offset)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx: Any, output_grad):
        data, scale, offset = ctx.saved_tensors
        dx, dscale, doffset = ops.quantize_by_tile_backward(data, output_grad, scale, ctx.tile_size, ctx.num_bits, offset)
        return dx, dscale, (doffset if offset is not None else None), None, None, None


class DequantizeAffine(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, data, scale, offset, tile_size, dtype):
        return ops.dequantize_by_tile(data, scale, _resolve_tile(data, tile_size), offset, dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx: Any, output_grad):
        return output_grad, None, None, None, None


class QuantizeDynamicAffine(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, data, tile_size, num_bits, symmetric, allow_one_sided, quantized_dtype):
        q, scale, offset = ops.quantize_dynamic_by_tile(
            data, _resolve_tile(data, tile_size), num_bits, symmetric, allow_one_sided, quantized_dtype or data.dtype)
        ctx.mark_non_differentiable(scale, offset)
        return q, scale, offset

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx: Any, output_grad, _scale_grad, _offset_grad):
        return output_grad, None, None, None, None, None


class FakeQuantizeAffine(torch.autograd.Function):
    """quantize -> dequantize in one kernel; backward == QuantizeStaticAffine.backward."""

    @staticmethod
    def forward(ctx: Any, data, scale, offset, tile_size, num_bits, quantized_dtype, dequantize_dtype):
        ctx.save_for_backward(data, scale, offset)
        ctx.tile_size = _resolve_tile(data, tile_size)
        ctx.num_bits = num_bits
        return ops.fake_quantize_by_tile(data, scale, ctx.tile_size, num_bits, quantized_dtype or data.dtype, offset,
                                         dequantize_dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx: Any, output_grad):
        data, scale, offset = ctx.saved_tensors
        dx, dscale, doffset = ops.quantize_by_tile_backward(data, output_grad, scale, ctx.tile_size, ctx.num_bits, offset)
        return dx, dscale, (doffset if offset is not None else None), None, None, None, None


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def quantize_affine(data, scale, offset, tile_size, num_bits: int, quantized_dtype: Optional[torch.dtype]):
    dt = _float_dtype(data)
    if not _needs_grad(data, scale, offset):      # inference / calibration: no autograd node to build
        return ops.quantize_by_tile(data, _tensor_or_none(scale, dt, data.device), _resolve_tile(data, tile_size), num_bits,
                                    quantized_dtype or data.dtype, _tensor_or_none(offset, dt, data.device))
    return QuantizeStaticAffine.apply(
        data, _tensor_or_none(scale, dt, data.device), _tensor_or_none(offset, dt, data.device), tile_size, num_bits,
        quantized_dtype)


def dequantize_affine(data, scale, offset, tile_size, dtype: Optional[torch.dtype]):
    if dtype is None:
        dtype = _float_dtype(data)
    if not _needs_grad(data, scale, offset):
        return ops.dequantize_by_tile(data, _tensor_or_none(scale, dtype, data.device), _resolve_tile(data, tile_size),
                                      _tensor_or_none(offset, dtype, data.device), dtype)
    return DequantizeAffine.apply(
        data, _tensor_or_none(scale, dtype, data.device), _tensor_or_none(offset, dtype, data.device), tile_size, dtype)


def quantize_dynamic_affine(data, tile_size, num_bits: int, symmetric: bool, allow_one_sided: bool, quantized_dtype):
    return QuantizeDynamicAffine.apply(data, tile_size, num_bits, symmetric, allow_one_sided, quantized_dtype)


def fake_quantize_affine(data, scale, offset, tile_size, num_bits: int, quantized_dtype, dequantize_dtype):
    dt = _float_dtype(data)
    return FakeQuantizeAffine.apply(
        data, _tensor_or_none(scale, dt, data.device), _tensor_or_none(offset, dt, data.device), tile_size, num_bits,
        quantized_dtype, dequantize_dtype)
