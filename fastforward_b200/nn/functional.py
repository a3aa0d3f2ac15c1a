"""Quantized functional operators (reference: _gen/operators.py:79-106, _gen/fallback.py:77-112).

Only ``linear`` is on the hot path.  It consults the dispatcher first -- that is where the W8A8
tensor-core kernel is registered (fastforward_b200/nn/qlinear.py) -- and otherwise takes the
reference's dequantize-then-float fallback."""

from __future__ import annotations

from typing import Optional

import torch

from .. import flags
from ..dispatcher import dispatch
from ..exceptions import QuantizationError
from ..quantized_tensor import QuantizedTensor


def _fallback_linear(input, weight, bias=None, *, output_quantizer=None, strict_quantization: bool = True):
    if strict_quantization and output_quantizer is None:
        raise QuantizationError("'output_quantizer' must be provided if strict_quantization=True")
    if strict_quantization and not isinstance(input, QuantizedTensor):
        raise QuantizationError("Expected 'input' to be an instance of 'QuantizedTensor' because strict_quantization=True.")
    if isinstance(input, QuantizedTensor):
        input = input.dequantize()
    if strict_quantization and not isinstance(weight, QuantizedTensor):
        raise QuantizationError("Expected 'weight' to be an instance of 'QuantizedTensor' because strict_quantization=True.")
    if isinstance(weight, QuantizedTensor):
        weight = weight.dequantize()
    if isinstance(bias, QuantizedTensor):
        bias = bias.dequantize()
    output = torch.nn.functional.linear(input, weight, bias)   # library float GEMM, as in the reference
    if output_quantizer is not None:
        output = output_quantizer(output)
    return output


class fallback:  # namespace mirroring fastforward._gen.fallback
    linear = staticmethod(_fallback_linear)


def linear(input: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
           output_quantizer=None, strict_quantization: Optional[bool] = None) -> torch.Tensor:
    if strict_quantization is None:
        strict_quantization = flags.get_strict_quantization()
    kwargs = dict(input=input, weight=weight, bias=bias, output_quantizer=output_quantizer,
                  strict_quantization=strict_quantization)
    op = dispatch("linear", **kwargs) or _fallback_linear
    return op(**kwargs)
