"""Functional front-ends for dynamic (per-call min/max) affine quantization
(reference: quantization/affine/dynamic.py:22-241)."""

from __future__ import annotations

from typing import Optional

import torch

from .. import granularity as granularities
from ..function import QuantizationContext
from .function import AffineQuantizationFunction, DynamicAffineQuantParams


def quantization_context(granularity=None, num_bits: int = 8, symmetric: bool = False, allow_one_sided: bool = True,
                         output_dtype: Optional[torch.dtype] = None, parameter_inference_fn=None) -> QuantizationContext:
    params = DynamicAffineQuantParams(
        num_bits=num_bits, granularity=granularity or granularities.PerTensor(), symmetric=symmetric,
        allow_one_sided=allow_one_sided, quantized_dtype=output_dtype, parameter_inference_fn=parameter_inference_fn)
    return QuantizationContext(AffineQuantizationFunction, params)


def quantize_per_granularity(input, granularity, num_bits: int = 8, output_dtype=None, symmetric: bool = False,
                             allow_one_sided: bool = True):
    ctx = quantization_context(granularity, num_bits, symmetric, allow_one_sided, output_dtype)
    return ctx.quantization_fn.quantize(input, ctx.quantization_params)


def quantize_by_tile(input, tile_size, num_bits: int = 8, output_dtype=None, symmetric: bool = False,
                     allow_one_sided: bool = True):
    return quantize_per_granularity(input, granularities.PerTile(tuple(tile_size)), num_bits, output_dtype, symmetric,
                                    allow_one_sided)


def quantize_per_tensor(input, num_bits: int = 8, output_dtype=None, symmetric: bool = False,
                        allow_one_sided: bool = True):
    return quantize_per_granularity(input, granularities.PerTensor(), num_bits, output_dtype, symmetric, allow_one_sided)


def quantize_per_channel(input, axis=-1, num_bits: int = 8, output_dtype=None, symmetric: bool = False,
                         allow_one_sided: bool = True):
    axes = (axis,) if isinstance(axis, int) else tuple(axis)
    axes = tuple(a % input.dim() for a in axes)
    return quantize_per_granularity(input, granularities.PerChannel(axes), num_bits, output_dtype, symmetric,
                                    allow_one_sided)


def quantize_per_block(input, channel_axis: int, block_axis: int, block_size: int, num_bits: int = 8,
                       output_dtype=None, symmetric: bool = False, allow_one_sided: bool = True):
    gran = granularities.PerBlock(block_dims=block_axis % input.dim(), block_sizes=block_size,
                                  per_channel_dims=channel_axis % input.dim())
    return quantize_per_granularity(input, gran, num_bits, output_dtype, symmetric, allow_one_sided)
