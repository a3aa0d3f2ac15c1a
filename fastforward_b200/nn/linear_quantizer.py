"""``LinearQuantizer``: affine quantizer with lazily materialised per-tile scale/offset
(reference: nn/linear_quantizer.py:25-357).

Differences that matter for B200: setting ``quantization_range`` runs ONE device kernel
(ops.parameters_for_range_) that writes scale and offset in place -- no host sync for the
one-sided decision, no temporaries -- and ``fake_quantize`` exposes the fused quantize->dequantize
pass.  Everything observable (parameter kinds, shapes, dtypes, error types, range getter) follows
the reference."""

from __future__ import annotations

import abc
from typing import Callable, Optional

import torch

from ..quantization import affine as affine_quant
from ..quantization import granularity as granularities
from ..quantization.function import QuantizationContext
from .quantizer import Quantizer

_Lazy = torch.nn.parameter.UninitializedTensorMixin


class AbstractAffineQuantizer(Quantizer, abc.ABC):
    def __init__(self, num_bits: int, *, granularity=None, quantized_dtype: Optional[torch.dtype] = None) -> None:
        super().__init__()
        self.num_bits = num_bits
        self.granularity = granularity or granularities.PerTensor()
        self.quantized_dtype = quantized_dtype

    per_channel = property(lambda self: granularities.is_per_channel(self.granularity))
    per_tensor = property(lambda self: granularities.is_per_tensor(self.granularity))
    integer_minimum = property(lambda self: affine_quant.integer_minimum(self.num_bits))
    integer_maximum = property(lambda self: affine_quant.integer_maximum(self.num_bits))

    @property
    def has_uninitialized_params(self) -> bool:
        # own parameters only (a quantizer has no submodules): this sits on every calibration step's host path
        for p in self._parameters.values():
            if isinstance(p, torch.nn.parameter.UninitializedParameter):
                return True
        return False

    def extra_repr(self) -> str:
        own = f"num_bits={self.num_bits}, granularity={self.granularity}"
        base = super().extra_repr()
        return f"{base}, {own}" if base else own

    @property
    @abc.abstractmethod
    def quantization_function(self):
        ...

    @abc.abstractmethod
    def quantization_parameters(self):
        ...

    def quantization_context(self) -> QuantizationContext:
        return QuantizationContext(self.quantization_function, self.quantization_parameters())

    def quantize(self, data: torch.Tensor) -> torch.Tensor:
        return self.quantization_function.quantize(data, self.quantization_parameters())


class LinearQuantizer(AbstractAffineQuantizer):
    def __init__(self, num_bits: int, *, symmetric: bool = True, allow_one_sided: bool = True, granularity=None,
                 quantized_dtype: Optional[torch.dtype] = None, param_dtype: Optional[torch.dtype] = None,
                 device="cpu") -> None:
        super().__init__(num_bits=num_bits, granularity=granularity, quantized_dtype=quantized_dtype)
        self.scale = torch.nn.UninitializedParameter(device=device, dtype=param_dtype)
        self.allow_one_sided = allow_one_sided
        if symmetric and not allow_one_sided:
            self.register_parameter("offset", None)
        elif symmetric:
            self.register_buffer("offset", torch.nn.UninitializedBuffer(device=device, dtype=param_dtype))
        else:
            self.offset = torch.nn.UninitializedParameter(device=device, dtype=param_dtype)

    @property
    def symmetric(self) -> bool:
        return "offset" in self._buffers or self.offset is None

    def reset_parameters(self) -> None:
        with torch.no_grad():
            self.scale = torch.nn.UninitializedParameter(device=self.scale.device, dtype=self.scale.dtype)
            if self.offset is not None:
                kind = torch.nn.UninitializedParameter if isinstance(self.offset, torch.nn.Parameter) \
                    else torch.nn.UninitializedBuffer
                self.offset = kind(device=self.offset.device, dtype=self.offset.dtype)

    def _initialize_parameters(self, parameter_dimensionality: int) -> None:
        if not self.has_uninitialized_params:
            return
        with torch.no_grad():
            self.scale.materialize((parameter_dimensionality,))
            self.scale.fill_(1.0)
            if self.offset is not None:
                self.offset.materialize((parameter_dimensionality,))
                self.offset.fill_(0.0)

    def extra_repr(self) -> str:
        own = f"symmetric={self.symmetric}"
        base = super().extra_repr()
        return f"{base}, {own}" if base else own

    def quantization_parameters(self) -> "affine_quant.StaticAffineQuantParams":
        return affine_quant.StaticAffineQuantParams(
            scale=self.scale, offset=self.offset, granularity=self.granularity, num_bits=self.num_bits,
            quantized_dtype=self.quantized_dtype)

    @property
    def quantization_function(self):
        return affine_quant.AffineQuantizationFunction

    def _uninitialized_error(self) -> ValueError:
        name = type(self).__name__
        return ValueError(
            f"Tried to quantize a tensor using an uninitialized quantizer (of type {name}). This quantizer is "
            f"initialized after its quantization_range is specified. This can be done explicitly by using the "
            f"{name}.quantization_range setter or using a range setting method.")

    def quantize(self, data: torch.Tensor) -> torch.Tensor:
        if isinstance(self.scale, _Lazy):
            raise self._uninitialized_error()
        return super().quantize(data)

    def fake_quantize(self, data: torch.Tensor) -> torch.Tensor:
        """``self(data).dequantize()`` without materialising the codes (one kernel)."""
        if isinstance(self.scale, _Lazy):
            raise self._uninitialized_error()
        return affine_quant.AffineQuantizationFunction.fake_quantize(data, self.quantization_parameters())

    def operator_for_range(self, min_range, max_range, data_shape) -> Callable[[torch.Tensor], torch.Tensor]:
        del data_shape
        scale, offset = self._parameters_for_range(min_range, max_range)
        ctx = affine_quant.quantization_context(scale=scale, offset=offset, num_bits=self.num_bits,
                                                granularity=self.granularity, output_dtype=self.quantized_dtype)
        return lambda data: ctx.quantization_fn.quantize(data, ctx.quantization_params)

    def _parameters_for_range(self, min_range, max_range):
        return affine_quant.parameters_for_range(min_range, max_range, self.num_bits, self.symmetric, self.allow_one_sided)

    @property
    def quantization_range(self):
        if self.has_uninitialized_params:
            return None, None
        return affine_quant.quantization_range(self.scale, self.offset, self.num_bits)

    @quantization_range.setter
    def quantization_range(self, quant_range) -> None:
        try:
            lo, hi = quant_range
        except ValueError as e:
            raise ValueError(f"Tried to set quantization range with {len(quant_range)}-tuple. A 2-tuple is expected") from e
        except TypeError as e:
            raise ValueError("Tried to set quantization range with a single value. A 2-tuple is expected") from e
        device = self.scale.device
        lo = lo if isinstance(lo, torch.Tensor) else torch.tensor(lo, device=device)
        hi = hi if isinstance(hi, torch.Tensor) else torch.tensor(hi, device=device)
        if self.has_uninitialized_params:
            self._initialize_parameters(lo.numel())
        self._set_range_(lo, hi)

    def _set_range_(self, lo: torch.Tensor, hi: torch.Tensor) -> None:
        """Write scale/offset for [lo, hi] in place with one kernel (linear_quantizer.py:350-357)."""
        from .. import ops

        if lo.numel() != self.scale.numel():
            if lo.numel() == 1 and hi.numel() == 1:
                lo, hi = lo.reshape(1).expand(self.scale.numel()), hi.reshape(1).expand(self.scale.numel())
            else:
                raise RuntimeError(
                    f"The size of the range ({lo.numel()}) must match the number of parameters ({self.scale.numel()})")
        if lo.device != self.scale.device:
            lo, hi = lo.to(self.scale.device), hi.to(self.scale.device)
        with torch.no_grad():
            ops.parameters_for_range_(lo, hi, self.num_bits, self.symmetric, self.allow_one_sided,
                                      self.scale.data, None if self.offset is None else self.offset.data)
