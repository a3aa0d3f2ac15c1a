"""CPU restatement of the calibration workload's hot path  --  TEST / BASELINE INFRASTRUCTURE ONLY.

``OracleCalibLinear`` is what ``QuantizedLinear.forward`` does in the reference while
``estimate_ranges(model, running_minmax)`` is active, written with the oracle's ops
(nn/linear.py:32-39 -> range_setting/common.py:218-238 -> range_setting/minmax.py:215-239 ->
nn/linear_quantizer.py:327-357 -> _gen/fallback.py:77-112), for W per-channel symmetric and
A per-tensor asymmetric quantizers.  bench.py's ``cpu_baseline`` and ``--impl reference`` legs time
it on the host cores; tests compare the CUDA path against it."""

from __future__ import annotations

from typing import Optional

import torch

from . import ref_ops as R


class OracleCalibLinear(torch.nn.Module):
    def __init__(self, linear: torch.nn.Linear, w_bits: int = 8, a_bits: int = 8, int8_codes: bool = True,
                 w_tile=None, disable_quantization: bool = False) -> None:
        super().__init__()
        self.weight, self.bias = linear.weight, linear.bias
        self.w_bits, self.a_bits = w_bits, a_bits
        self.qdtype = torch.int8 if int8_codes else None
        self.w_tile = w_tile
        self.disable_quantization = disable_quantization
        self.x_min = self.x_max = self.w_min = self.w_max = None
        self.x_scale = self.x_offset = self.w_scale = self.w_offset = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            x_tile = tuple(x.shape)                                        # PerTensor
            w_tile = self.w_tile or (1, self.weight.shape[1])              # PerChannel(0)
            self.x_min, self.x_max, self.x_scale, self.x_offset, xq = R.calibration_quantizer_step(
                self.x_min, self.x_max, x, x_tile, self.a_bits, False, True, self.qdtype)
            self.w_min, self.w_max, self.w_scale, self.w_offset, wq = R.calibration_quantizer_step(
                self.w_min, self.w_max, self.weight, w_tile, self.w_bits, True, True, self.qdtype)
            if self.disable_quantization:
                return torch.nn.functional.linear(x, self.weight, self.bias)
            return R.fallback_linear(xq, self.x_scale, self.x_offset, x_tile, x.dtype,
                                     wq, self.w_scale, self.w_offset, w_tile, self.weight.dtype, self.bias)


def oracle_calibration_model(model: torch.nn.Module, **kwargs) -> torch.nn.Module:
    """Swap every nn.Linear under ``model.layers`` for its oracle calibration counterpart (in place)."""
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            if isinstance(child, torch.nn.Linear):
                setattr(parent, name, OracleCalibLinear(child, **kwargs))
    return model
