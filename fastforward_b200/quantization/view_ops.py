"""Element-rearranging ops on per-tensor affine ``QuantizedTensor``s (reference:
quantization/_linear_quantized_ops.py:92-230): with ONE (scale, offset) for the whole tensor, an op that only moves
elements around -- view / reshape / transpose / permute / unsqueeze / expand / indexing -- is the same op on the raw
codes under the same quantization context.  No arithmetic, no kernels: what these buy is that the attention matmuls
(``q @ k.transpose(-1, -2)``) keep their int8 operands and reach the tcgen05 kernels registered in nn/qlinear.py
instead of being dequantized implicitly."""

from __future__ import annotations

from typing import Any

import torch

from ..dispatcher import Predicate, register
from ..quantized_tensor import QuantizedTensor
from . import granularity as G


def _is_affine_per_tensor(input: Any = None, *_a: Any, **_k: Any) -> bool:
    if not isinstance(input, QuantizedTensor):
        return False
    from .affine.function import AffineQuantizationFunction, StaticAffineQuantParams

    ctx = input.quantization_context
    return (isinstance(ctx.quantization_fn, type) and issubclass(ctx.quantization_fn, AffineQuantizationFunction)
            and isinstance(ctx.quantization_params, StaticAffineQuantParams)
            and G.is_per_tensor(ctx.quantization_params.granularity))


affine_per_tensor_predicate = Predicate(_is_affine_per_tensor)


def _rearranging(name: str):
    tensor_fn = getattr(torch.Tensor, name)

    def op(input: QuantizedTensor, *args: Any, **kwargs: Any) -> QuantizedTensor:
        if name in ("view", "view_as") and args and isinstance(args[0], torch.dtype):
            raise TypeError(f"QuantizedTensor.{name}(dtype) is not supported")
        raw = tensor_fn(input.raw_data, *args, **kwargs)
        return input.quantization_context.attach(raw)

    op.__name__ = name
    return op


for _name in ("view", "view_as", "reshape", "transpose", "permute", "unsqueeze", "squeeze", "expand", "flatten",
              "__getitem__", "t"):
    register(_name, affine_per_tensor_predicate, _rearranging(_name))
