"""``DynamicLinearQuantizer``: parameters inferred from every input (per-tile min/max) instead of
stored (reference: nn/dynamic_linear_quantizer.py:20-92).  One call = min/max kernel -> range->params
kernel -> quantize kernel, no host sync (ffq_dynamic_quantize)."""

from __future__ import annotations

from typing import Optional

import torch

from ..quantization import affine as affine_quant
from .linear_quantizer import AbstractAffineQuantizer


class DynamicLinearQuantizer(AbstractAffineQuantizer):
    def __init__(self, num_bits: int, *, granularity=None, quantized_dtype: Optional[torch.dtype] = None,
                 parameter_inference_fn=None, allow_one_sided: bool = True, symmetric: bool = False) -> None:
        super().__init__(num_bits=num_bits, granularity=granularity, quantized_dtype=quantized_dtype)
        self.parameter_inference_fn = parameter_inference_fn
        self.symmetric = symmetric
        self.allow_one_sided = allow_one_sided

    def quantization_parameters(self) -> "affine_quant.DynamicAffineQuantParams":
        return affine_quant.DynamicAffineQuantParams(
            granularity=self.granularity, num_bits=self.num_bits, quantized_dtype=self.quantized_dtype,
            parameter_inference_fn=self.parameter_inference_fn, symmetric=self.symmetric,
            allow_one_sided=self.allow_one_sided)

    @property
    def quantization_function(self):
        return affine_quant.AffineQuantizationFunction

    def reset_parameters(self) -> None:
        pass
