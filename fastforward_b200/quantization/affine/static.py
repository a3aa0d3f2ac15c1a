"""Functional front-ends for static affine quantization
(reference: quantization/affine/static.py:19-213)."""

from __future__ import annotations

from typing import Optional

import torch

from .. import granularity as granularities
from ..function import QuantizationContext
from .function import AffineQuantizationFunction, StaticAffineQuantParams


def quantization_context(scale, offset, granularity=None, num_bits: int = 8,
                         output_dtype: Optional[torch.dtype] = None,
                         dequantize_dtype: Optional[torch.dtype] = None) -> QuantizationContext:
    params = StaticAffineQuantParams(
        scale=scale, offset=offset, num_bits=num_bits, granularity=granularity or granularities.PerTensor(),
        quantized_dtype=output_dtype, dequantize_dtype=dequantize_dtype)
    return QuantizationContext(AffineQuantizationFunction, params)


def quantize_per_granularity(input, scale, offset, granularity, num_bits: int = 8, output_dtype=None):
    ctx = quantization_context(scale, offset, granularity, num_bits, output_dtype)
    return ctx.quantization_fn.quantize(input, ctx.quantization_params)


def quantize_by_tile(input, scale, offset, tile_size, num_bits: int = 8, output_dtype=None):
    return quantize_per_granularity(input, scale, offset, granularities.PerTile(tuple(tile_size)), num_bits, output_dtype)


def quantize_per_tensor(input, scale, offset=None, num_bits: int = 8, output_dtype=None):
    return quantize_per_granularity(input, scale, offset, granularities.PerTensor(), num_bits, output_dtype)


def quantize_per_channel(input, scale, offset=None, axis=-1, num_bits: int = 8, output_dtype=None):
    axes = (axis,) if isinstance(axis, int) else tuple(axis)
    axes = tuple(a % input.dim() for a in axes)
    return quantize_per_granularity(input, scale, offset, granularities.PerChannel(axes), num_bits, output_dtype)


def quantize_per_block(input, scale, offset, channel_axis: int, block_axis: int, block_size: int,
                       num_bits: int = 8, output_dtype=None):
    gran = granularities.PerBlock(block_dims=block_axis % input.dim(), block_sizes=block_size,
                                  per_channel_dims=channel_axis % input.dim())
    return quantize_per_granularity(input, scale, offset, gran, num_bits, output_dtype)
