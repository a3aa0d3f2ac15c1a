"""Minimum-error (MSE grid search) range estimator (reference: range_setting/min_error.py:64-317).

Behaviour is the reference's, including two quirks callers can observe:
  * the search grid is built once, from the FIRST batch (`setup_estimator`, :188-195);
  * the estimator loops ``range(num_candidates)`` (:207) although the asymmetric grid has
    ``floor(sqrt(C)) * (floor(sqrt(C)) + C - floor(sqrt(C))**2) >= C`` rows (:121-139): rows past
    ``num_candidates`` keep an accumulated error of 0 and therefore win the arg-min.

What changed for B200 is the schedule.  The reference runs quantize -> dequantize -> (y-x)^2 ->
mean once per candidate: C x ~30 passes over the tensor per batch.  Here the candidates'
(scale, offset) are computed once when the grid is built and ONE kernel (`ops.grid_mse`) reads
the batch once and evaluates every candidate from registers.  That fused path is taken when the
error function is `mse_error`, the quantizer is a `LinearQuantizer` and the tiles are contiguous
runs; anything else evaluates candidate by candidate through `operator_for_range`, with the same
CUDA ops underneath.
"""

from __future__ import annotations

import dataclasses
import logging
from math import floor, sqrt
from typing import Callable, Iterator, Optional, Protocol, runtime_checkable

import torch

from .. import ops
from ..nn.quantized_module import named_quantizers
from ..nn.quantizer import Quantizer
from ..quantization.tiled_tensor import tiles_to_rows
from .common import RangeEstimator, RangeSettable, SimpleEstimatorStep

logger = logging.getLogger(__name__)


@runtime_checkable
class SupportsRangeBasedOperator(RangeSettable, Protocol):
    """Quantizers that can build a quantization operator for a given range (common.py:69-107)."""

    @property
    def symmetric(self) -> bool: ...

    def operator_for_range(self, __min, __max, __data_shape): ...


def mse_error(quantized_data: torch.Tensor, unquantized_data: torch.Tensor) -> torch.Tensor:
    """Row-wise mean squared error (min_error.py:64-74)."""
    return torch.mean((quantized_data - unquantized_data) ** 2, dim=1)


def _host_linspace(start: float, end: float, steps: int, like: torch.Tensor) -> torch.Tensor:
    # C values, computed where the reference's golden vectors pin them (aten's CPU linspace), then moved
    return torch.linspace(start, end, steps, dtype=like.dtype).to(like.device)


@dataclasses.dataclass
class _UniformSearchGrid:
    absolute_margin: float = 0.5
    relative_margin: float = 1.0

    def __call__(self, tiled_data_sample: torch.Tensor, symmetric: bool, parameter_dimensionality: int,
                 num_candidates: int):
        """(min_threshold, max_threshold), each ``[num_candidates, parameter_dimensionality]``
        (``[n_lo * n_hi, ...]`` on the asymmetric branch), from a ``[tiles, tile_numel]`` sample."""
        assert tiled_data_sample.ndim == 2
        assert tiled_data_sample.shape[0] == parameter_dimensionality
        lo, hi = ops.tile_minmax(tiled_data_sample, (1, tiled_data_sample.shape[1]))
        return self.from_minmax(lo, hi, symmetric, num_candidates)

    def from_minmax(self, tile_min: torch.Tensor, tile_max: torch.Tensor, symmetric: bool, num_candidates: int):
        """The same grid from per-tile extrema (min_error.py:102-146) -- the row view of the data is
        never materialised."""
        rel, absm = self.relative_margin, self.absolute_margin
        max_data = rel * tile_max + absm
        min_data = rel * tile_min - absm
        negative_data = bool(min_data.min() < 0)           # one host sync, at grid construction only
        if not negative_data:
            min_threshold = torch.zeros((num_candidates, tile_min.numel()), dtype=tile_min.dtype, device=tile_min.device)
            steps = _host_linspace(1 / num_candidates, 1, num_candidates, tile_min)
            max_threshold = steps.unsqueeze(1) * max_data.unsqueeze(0)
        elif not symmetric:
            margin = 0.6
            n_lo = floor(sqrt(num_candidates))
            n_hi = n_lo + num_candidates - n_lo ** 2
            steps_lo = _host_linspace(1, margin, n_lo, tile_min)
            steps_hi = _host_linspace(margin, 1, n_hi, tile_min)
            min_threshold = steps_lo.unsqueeze(1) * (rel * min_data.unsqueeze(0) + absm)
            max_threshold = steps_hi.unsqueeze(1) * (rel * max_data.unsqueeze(0) + absm)
            min_threshold = min_threshold.repeat(n_hi, 1)
            max_threshold = max_threshold.repeat_interleave(n_lo, dim=0)
        else:
            steps = _host_linspace(1 / num_candidates, 1, num_candidates, tile_min)
            max_abs = torch.max(torch.abs(min_data), torch.abs(max_data))
            max_threshold = steps.unsqueeze(1) * max_abs.unsqueeze(0)
            min_threshold = -max_threshold
        return min_threshold, max_threshold


def uniform_search_grid(absolute_margin: float = 0.5, relative_margin: float = 1.0) -> _UniformSearchGrid:
    """Search ranges within ``(r * min - a, r * max + a)`` of the first batch (min_error.py:151-170)."""
    return _UniformSearchGrid(absolute_margin=absolute_margin, relative_margin=relative_margin)


class _MinAvgErrorGridEstimator(SimpleEstimatorStep, torch.nn.Module):
    def __init__(self, quantizer, error_fn: Callable = mse_error, num_candidates: int = 100,
                 search_grid_generator: Callable = _UniformSearchGrid(),
                 update_range_policy: Optional[Callable[["_MinAvgErrorGridEstimator", int], bool]] = None,
                 disable_quantization: bool = False) -> None:
        super().__init__(disable_quantization=disable_quantization)
        self._quantizer = quantizer
        self.error_fn = error_fn
        self.num_candidates = num_candidates
        self.search_grid_generator = search_grid_generator
        self._estimation_steps = 0
        self.update_range_policy = update_range_policy
        self._cand_scale = self._cand_offset = None
        self._fused = None           # None = undecided, True/False once the first step has run

    # ---- grid -------------------------------------------------------------------------------
    def setup_estimator(self, data: torch.Tensor) -> None:
        self._estimation_steps = 0
        self._initialize_search_grid(data)

    def _initialize_search_grid(self, data: torch.Tensor) -> None:
        granularity = self._quantizer.granularity
        n_params = granularity.parameter_dimensionality(data.shape)
        tile = granularity.tile_size(data.shape)
        gen = self.search_grid_generator
        with torch.no_grad():
            if isinstance(gen, _UniformSearchGrid):
                lo, hi = ops.tile_minmax(data.detach(), tile)
                self.min_threshold, self.max_threshold = gen.from_minmax(
                    lo, hi, self._quantizer.symmetric, self.num_candidates)
            else:
                self.min_threshold, self.max_threshold = gen(
                    tiles_to_rows(data.detach(), tile), self._quantizer.symmetric, n_params, self.num_candidates)
        self.cumulative_error = torch.zeros_like(self.min_threshold)
        self._cand_scale = self._cand_offset = None
        self._fused = None

    def _fusable(self, quantizer, data: torch.Tensor) -> bool:
        from ..nn.linear_quantizer import LinearQuantizer

        return (self.error_fn is mse_error and type(quantizer).operator_for_range is LinearQuantizer.operator_for_range
                and data.is_cuda and data.dtype in (torch.float32, torch.float16, torch.bfloat16))

    def _candidate_parameters(self, quantizer) -> None:
        """(scale, offset) of the candidates that will be evaluated, ``[C, tiles]`` fp32 -- once per
        grid.  Each candidate takes its own one-sided decision, as `operator_for_range` does."""
        n = min(self.num_candidates, self.min_threshold.shape[0])
        tiles = self.min_threshold.shape[1]
        dev = self.min_threshold.device
        self._cand_scale = torch.empty((n, tiles), dtype=torch.float32, device=dev)
        has_offset = not (quantizer.symmetric and not quantizer.allow_one_sided)
        self._cand_offset = torch.empty((n, tiles), dtype=torch.float32, device=dev) if has_offset else None
        for i in range(n):
            ops.parameters_for_range_(
                self.min_threshold[i], self.max_threshold[i], quantizer.num_bits, quantizer.symmetric,
                quantizer.allow_one_sided, self._cand_scale[i], None if not has_offset else self._cand_offset[i])

    # ---- per batch ----------------------------------------------------------------------------
    def _update_quantizer_ranges(self, quantizer) -> None:
        best = self.cumulative_error.min(dim=0).indices
        idx = torch.arange(self.min_threshold.shape[1], device=best.device)
        quantizer.quantization_range = (self.min_threshold[best, idx], self.max_threshold[best, idx])

    def estimate_step(self, quantizer, data: torch.Tensor) -> None:
        tile = self._quantizer.granularity.tile_size(data.shape)
        with torch.no_grad():
            x = data.detach()
            if self._fused is None:
                self._fused = self._fusable(quantizer, x)
            if self._fused:
                if self._cand_scale is None:
                    self._candidate_parameters(quantizer)
                try:
                    err = ops.grid_mse(x, self._cand_scale, self._cand_offset, tile, quantizer.num_bits,
                                       quantizer.quantized_dtype)
                    self.cumulative_error[: err.shape[0]] += err.to(self.cumulative_error.dtype)
                except NotImplementedError:      # tiles are not contiguous runs: candidate by candidate
                    self._fused = False
            if not self._fused:
                rows = tiles_to_rows(x, tile)
                for i in range(self.num_candidates):
                    quant_op = quantizer.operator_for_range(self.min_threshold[i], self.max_threshold[i], data.shape)
                    quant_data = quant_op(x).dequantize()
                    self.cumulative_error[i] += self.error_fn(tiles_to_rows(quant_data, tile), rows)
        self._estimation_steps += 1
        if not self.update_range_policy or self.update_range_policy(self, self._estimation_steps):
            self._update_quantizer_ranges(quantizer)


class MinErrorGridRangeEstimator(RangeEstimator):
    """Grid search for the range minimising ``error_fn(quantized, original)`` (min_error.py:224-315)."""

    def __init__(self, error_fn: Callable = mse_error, num_candidates: int = 100,
                 search_grid_generator: Callable = _UniformSearchGrid(),
                 update_range_policy: Optional[Callable] = None, skip_unsupported_quantizers: bool = False) -> None:
        self._error_fn = error_fn
        self._num_candidates = num_candidates
        self._search_grid_generator = search_grid_generator
        self._update_range_policy = update_range_policy
        self._skip_unsupported_quantizers = skip_unsupported_quantizers

    def prepare(self, module):
        if not isinstance(module, SupportsRangeBasedOperator):
            proto = f"{SupportsRangeBasedOperator.__module__}.{SupportsRangeBasedOperator.__qualname__}"
            raise TypeError(f"{type(module).__name__} does not implement {proto}.")
        return module.register_override(_MinAvgErrorGridEstimator(
            module, error_fn=self._error_fn, num_candidates=self._num_candidates,
            search_grid_generator=self._search_grid_generator, update_range_policy=self._update_range_policy))

    def cleanup(self, module, metadata) -> None:
        del module
        metadata.remove()

    def split_module(self, module: torch.nn.Module) -> Iterator[Quantizer]:
        quantizers = [("", module)] if isinstance(module, Quantizer) and not module.is_stub() else []
        quantizers += [(n, q) for n, q in named_quantizers(module, recurse=True) if q is not module]
        for _, quantizer in quantizers:
            if isinstance(quantizer, SupportsRangeBasedOperator) or not self._skip_unsupported_quantizers:
                yield quantizer
            else:
                logger.warning(f"{type(quantizer).__name__} does not implement SupportsRangeBasedOperator. "
                               f"Therefore it is not included in {type(self).__name__} range setting.")


min_error_grid = MinErrorGridRangeEstimator
mse_grid = MinErrorGridRangeEstimator
