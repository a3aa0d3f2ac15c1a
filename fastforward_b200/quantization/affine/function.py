"""``AffineQuantizationFunction`` and its parameter records
(reference: quantization/affine/function.py:31-212)."""

from __future__ import annotations

import dataclasses
from typing import Any, Callable, Optional

import torch

from ... import flags
from ...exceptions import ExportError
from .. import granularity as granularities
from ..function import QuantizationContext, QuantizationFunction, QuantizationParameters
from ._autograd import dequantize_affine, fake_quantize_affine, quantize_affine, quantize_dynamic_affine


@dataclasses.dataclass
class StaticAffineQuantParams(QuantizationParameters):
    scale: Any
    offset: Any
    num_bits: int
    granularity: granularities.Granularity
    quantized_dtype: Optional[torch.dtype] = None
    dequantize_dtype: Optional[torch.dtype] = None


@dataclasses.dataclass
class DynamicAffineQuantParams(QuantizationParameters):
    num_bits: int
    granularity: granularities.Granularity
    symmetric: bool = False
    allow_one_sided: bool = True
    quantized_dtype: Optional[torch.dtype] = None
    dequantize_dtype: Optional[torch.dtype] = None
    parameter_inference_fn: Optional[Callable[..., Any]] = None


DynamicParamInferenceFn = Callable[[DynamicAffineQuantParams, torch.Tensor], Any]


def _static_from_dynamic(params: DynamicAffineQuantParams, scale, offset, **changes: Any) -> StaticAffineQuantParams:
    keep = {f.name for f in dataclasses.fields(StaticAffineQuantParams)}
    args = {k: v for k, v in params._fields().items() if k in keep}
    args.update(scale=scale, offset=offset, **changes)
    return StaticAffineQuantParams(**args)


class AffineQuantizationFunction(QuantizationFunction):
    @classmethod
    def quantize(cls, data: torch.Tensor, params):
        if flags.get_export_mode():
            return cls._export_quantize(data, params)
        if isinstance(params, StaticAffineQuantParams):
            return cls._static_quantize(data, params)
        if isinstance(params, DynamicAffineQuantParams):
            return cls._dynamic_quantize(data, params)
        raise TypeError(f"Unsupported type for argument 'params': '{type(params)}'")

    @classmethod
    def _export_quantize(cls, data: torch.Tensor, params) -> torch.Tensor:
        """Quantize immediately followed by dequantize, returning a plain tensor
        (function.py:94-121) -- here a single fused kernel."""
        if not isinstance(params, StaticAffineQuantParams):
            raise ExportError("Export supports only static affine quantization.")
        tile_size = params.granularity.tile_size(data.shape)
        qdtype = params.quantized_dtype or data.dtype
        return fake_quantize_affine(data, params.scale, params.offset, tile_size, params.num_bits, qdtype, qdtype)

    @classmethod
    def fake_quantize(cls, data: torch.Tensor, params: StaticAffineQuantParams) -> torch.Tensor:
        """``quantize(data).dequantize()`` in one pass (same bits, same gradients)."""
        tile_size = params.granularity.tile_size(data.shape)
        return fake_quantize_affine(data, params.scale, params.offset, tile_size, params.num_bits,
                                    params.quantized_dtype or data.dtype, params.dequantize_dtype or data.dtype)

    @classmethod
    def _static_quantize(cls, data: torch.Tensor, params: StaticAffineQuantParams):
        tile_size = params.granularity.tile_size(data.shape)
        raw = quantize_affine(data, params.scale, params.offset, tile_size, params.num_bits,
                              params.quantized_dtype or data.dtype)
        params = params.with_changes(dequantize_dtype=params.dequantize_dtype or data.dtype)
        from ...quantized_tensor import QuantizedTensor

        return QuantizedTensor(raw, QuantizationContext(cls, params))

    @classmethod
    def _dynamic_quantize(cls, data: torch.Tensor, params: DynamicAffineQuantParams):
        if params.parameter_inference_fn is not None:
            scale, offset = params.parameter_inference_fn(params, data)
            return cls._static_quantize(
                data, _static_from_dynamic(params, scale, offset, dequantize_dtype=params.dequantize_dtype or data.dtype))
        tile_size = params.granularity.tile_size(data.shape)
        tile_size = data.shape if isinstance(tile_size, str) else tile_size
        raw, scale, offset = quantize_dynamic_affine(
            data, tile_size, params.num_bits, params.symmetric, params.allow_one_sided,
            params.quantized_dtype or data.dtype)
        static = _static_from_dynamic(params, scale, offset, dequantize_dtype=params.dequantize_dtype or data.dtype)
        from ...quantized_tensor import QuantizedTensor

        return QuantizedTensor(raw, QuantizationContext(AffineQuantizationFunction, static))

    @classmethod
    def dequantize(cls, data: torch.Tensor, params) -> torch.Tensor:
        if isinstance(params, DynamicAffineQuantParams):
            raise TypeError("Cannot dequantize a QuantizedTensor with dynamic parameters.")
        tile_size = params.granularity.tile_size(data.shape)
        return dequantize_affine(data, params.scale, params.offset, tile_size, params.dequantize_dtype)
