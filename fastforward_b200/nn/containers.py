"""Quantized counterparts of the container / trivially-wrapped modules a Llama-style model needs
(reference: nn/container.py, nn/activations.py, nn/embedding.py, nn/normalization.py -- same
pattern as linear; only what the calibration workload touches is provided here)."""

from __future__ import annotations

import torch

from ..quantized_tensor import QuantizedTensor
from .quantized_module import QuantizedModule
from .quantizer import QuantizerStub


class QuantizedSequential(QuantizedModule, torch.nn.Sequential):
    pass


class QuantizedModuleList(QuantizedModule, torch.nn.ModuleList):
    pass


class QuantizedModuleDict(QuantizedModule, torch.nn.ModuleDict):
    pass


class QuantizedSiLU(QuantizedModule, torch.nn.SiLU):
    def __init_quantization__(self) -> None:
        super().__init_quantization__()
        self.output_quantizer = QuantizerStub(output_quantizer=True)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        x = input.dequantize() if isinstance(input, QuantizedTensor) else input
        return self.output_quantizer(torch.nn.functional.silu(x))


class QuantizedEmbedding(QuantizedModule, torch.nn.Embedding):
    def __init_quantization__(self) -> None:
        super().__init_quantization__()
        self.weight_quantizer = QuantizerStub(weight_quantizer=True, shape=self.weight.shape)
        self.output_quantizer = QuantizerStub(output_quantizer=True)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        weight = self.weight_quantizer(self.weight)
        weight = weight.dequantize() if isinstance(weight, QuantizedTensor) else weight
        out = torch.nn.functional.embedding(input, weight, self.padding_idx, self.max_norm, self.norm_type,
                                            self.scale_grad_by_freq, self.sparse)
        return self.output_quantizer(out)
