"""Quantized functional operators (reference: _gen/operators.py:79-106, _gen/fallback.py:77-112).

``linear`` and the attention matmuls (``matmul`` / ``mm`` / ``bmm``) are on the hot path.  Each consults the
dispatcher first -- that is where the tensor-core kernels are registered (fastforward_b200/nn/qlinear.py) -- and
otherwise takes the reference's dequantize-then-float fallback."""

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


def _fallback_binary(torch_fn, second: str):
    """The reference's generated fallback for matmul / mm / bmm (_gen/fallback.py:699-800): strict checks,
    dequantize both operands, the float library op, then the output quantizer."""

    def fallback_op(input, other=None, *, output_quantizer=None, strict_quantization: bool = True, **kw):
        if other is None:
            other = kw.pop(second)
        if strict_quantization and output_quantizer is None:
            raise QuantizationError("'output_quantizer' must be provided if strict_quantization=True")
        if strict_quantization and not isinstance(input, QuantizedTensor):
            raise QuantizationError("Expected 'input' to be an instance of 'QuantizedTensor' because strict_quantization=True.")
        if isinstance(input, QuantizedTensor):
            input = input.dequantize()
        if strict_quantization and not isinstance(other, QuantizedTensor):
            raise QuantizationError(f"Expected '{second}' to be an instance of 'QuantizedTensor' because strict_quantization=True.")
        if isinstance(other, QuantizedTensor):
            other = other.dequantize()
        output = torch_fn(input, other)
        if output_quantizer is not None:
            output = output_quantizer(output)
        return output

    return fallback_op


class fallback:  # namespace mirroring fastforward._gen.fallback
    linear = staticmethod(_fallback_linear)
    matmul = staticmethod(_fallback_binary(torch.matmul, "other"))
    mm = staticmethod(_fallback_binary(torch.mm, "mat2"))
    bmm = staticmethod(_fallback_binary(torch.bmm, "mat2"))


def linear(input: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
           output_quantizer=None, strict_quantization: Optional[bool] = None) -> torch.Tensor:
    if strict_quantization is None:
        strict_quantization = flags.get_strict_quantization()
    kwargs = dict(input=input, weight=weight, bias=bias, output_quantizer=output_quantizer,
                  strict_quantization=strict_quantization)
    op = dispatch("linear", **kwargs) or _fallback_linear
    return op(**kwargs)


def _binary(name: str, second: str):
    def op(input: torch.Tensor, other: torch.Tensor, *, output_quantizer=None,
           strict_quantization: Optional[bool] = None) -> torch.Tensor:
        if strict_quantization is None:
            strict_quantization = flags.get_strict_quantization()
        kwargs = {"input": input, second: other, "output_quantizer": output_quantizer,
                  "strict_quantization": strict_quantization}
        kernel = dispatch(name, **kwargs) or getattr(fallback, name)
        return kernel(**kwargs)

    op.__name__ = name
    op.__doc__ = f"Quantized ``torch.{name}`` (reference: _gen/operators.py:654-735): dispatcher first, then the fallback."
    return op


matmul = _binary("matmul", "other")
mm = _binary("mm", "mat2")
bmm = _binary("bmm", "mat2")
