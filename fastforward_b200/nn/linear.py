"""``QuantizedLinear`` (reference: nn/linear.py:12-39): four quantizer slots around ``linear``."""

from __future__ import annotations

import torch

from .functional import linear
from .quantized_module import QuantizedModule
from .quantizer import QuantizerStub


class QuantizedLinear(QuantizedModule, torch.nn.Linear):
    def __init_quantization__(self) -> None:
        super().__init_quantization__()
        self.input_quantizer = QuantizerStub(input_quantizer=True)
        self.weight_quantizer = QuantizerStub(weight_quantizer=True, shape=self.weight.shape)
        if self.bias is not None:
            self.bias_quantizer = QuantizerStub(bias_quantizer=True, shape=self.bias.shape)
        else:
            self.register_quantizer("bias_quantizer", None)
        self.output_quantizer = QuantizerStub(output_quantizer=True)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        input = self.input_quantizer(input)
        weight = self.weight_quantizer(self.weight)
        bias = self.bias
        if bias is not None and self.bias_quantizer is not None:
            bias = self.bias_quantizer(bias)
        return linear(input, weight, bias, output_quantizer=self.output_quantizer)
