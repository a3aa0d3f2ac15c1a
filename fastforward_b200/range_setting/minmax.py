"""Running / smoothed min-max range estimators (reference: range_setting/minmax.py:26-303).

What changed for B200 -- behaviour is the same, the schedule is not:
  * one fused kernel reads the tensor ONCE and merges per-tile min/max into the running range
    (the reference reads it twice and allocates two temporaries);
  * the +-inf check (minmax.py:233-234, a host sync per quantizer per forward) is a device flag
    inspected once when the ``estimate_ranges`` block ends -- pass ``eager_checks=True`` to get
    the reference's raise-immediately behaviour back (one sync per step);
  * range -> (scale, offset) is one more kernel writing the quantizer's parameters in place, with
    the one-sided decision taken on the device (no second host sync);
  * ``process_group`` (or an initialised default group with ``sync_ranges=True``) all-reduces the
    running ranges with MIN/MAX at the end of the block: data-parallel calibration;
  * when the step is followed by the quantizer's own int8 quantization (the usual W8A8 calibration
    forward) and the layout is a per-channel row or a whole tensor, estimate_step + range setter +
    quantize (+ the code row sums the W8A8 linear needs) are ONE kernel reading the tensor once
    (``ops.calibrate_quantize_``; pass ``fused=False`` to the estimator to keep the separate kernels).
"""

from __future__ import annotations

import logging
from typing import Iterator, Optional, Sequence

import torch

from .. import flags as _flags
from .. import ops
from ..forward_override import _Chain
from ..nn.quantized_module import named_quantizers
from ..nn.quantizer import Quantizer
from .common import RangeEstimator, RangeSettable, SimpleEstimatorStep

logger = logging.getLogger(__name__)


def _param_count(quantizer, data: torch.Tensor) -> int:
    return quantizer.granularity.parameter_dimensionality(data.shape)


class _RangeArena:
    """Running ranges of ALL quantizers of one estimate_ranges block live in a few flat buffers
    (one min and one max buffer per dtype/device), so the data-parallel exchange at the end of the
    block is two collectives on contiguous memory -- no packing of hundreds of tiny tensors -- and
    one flag word serves every quantizer."""

    CHUNK = 1 << 21

    def __init__(self) -> None:
        self.chunks: dict = {}     # (device, dtype) -> list of [min_buf, max_buf, used]
        self.flags: dict = {}      # device -> int32[1]
        self.barriers: dict = {}   # device -> zeroed scratch of the fused per-tensor kernel's grid barrier

    def take(self, n: int, like: torch.Tensor):
        key = (like.device, like.dtype)
        chunks = self.chunks.setdefault(key, [])
        if not chunks or chunks[-1][2] + n > chunks[-1][0].numel():
            size = max(self.CHUNK, n)
            chunks.append([like.new_full((size,), float("inf")), like.new_full((size,), float("-inf")), 0])
        mn, mx, used = chunks[-1]
        chunks[-1][2] = used + n
        return mn[used:used + n], mx[used:used + n]

    def flag(self, device: torch.device) -> torch.Tensor:
        if device not in self.flags:
            self.flags[device] = torch.zeros(1, dtype=torch.int32, device=device)
        return self.flags[device]

    def barrier_workspace(self, device: torch.device) -> torch.Tensor:
        """Owned by the block (like the running ranges), so a CUDA graph captured inside the block never points at
        scratch that outlives it; one stream per block is assumed, as everywhere in the reference."""
        if device not in self.barriers:
            self.barriers[device] = torch.zeros(ops._CALQ_WS, dtype=torch.uint8, device=device)
        return self.barriers[device]

    def buffers(self):
        for chunks in self.chunks.values():
            for mn, mx, used in chunks:
                if used:
                    yield mn[:used], mx[:used]


class RunningMinMaxEstimator(SimpleEstimatorStep, torch.nn.Module):
    def __init__(self, quantizer, disable_quantization: bool = False, eager_checks: bool = False,
                 arena: Optional[_RangeArena] = None, fused: bool = True) -> None:
        super().__init__(disable_quantization=disable_quantization)
        self._arena = arena
        self._fused = fused
        self._settled = None          # device int32[1]: set by the kernel once every running min is negative
        self._settled_host = None     # pinned mirror, refreshed asynchronously: lets later steps skip the fix-up launch
        self._settled_seen = False
        lo, hi = quantizer.quantization_range       # continues from an existing range (minmax.py:198-200)
        self.register_buffer("min", None if lo is None else lo.detach().clone())
        self.register_buffer("max", None if hi is None else hi.detach().clone())
        self.register_buffer("flags", None)
        self._eager = eager_checks

    def initialize_parameters(self, quantizer, data: torch.Tensor) -> None:
        n = _param_count(quantizer, data)
        if self.min is None and self.max is None and self._arena is not None:
            self.min, self.max = self._arena.take(n, data)
        if self.min is None:
            self.min = data.new_full((n,), float("inf"))
        if self.max is None:
            self.max = data.new_full((n,), float("-inf"))
        if self.min.dtype != data.dtype or self.min.device != data.device:
            # torch.min(self.min, data_min) in the reference promotes; keep the running range in the
            # promoted dtype so that nothing is lost
            dt = torch.promote_types(self.min.dtype, data.dtype)
            self.min, self.max = self.min.to(device=data.device, dtype=dt), self.max.to(device=data.device, dtype=dt)
        if self.flags is None:
            self.flags = self._arena.flag(data.device) if self._arena is not None else \
                torch.zeros(1, dtype=torch.int32, device=data.device)

    def estimate_step(self, quantizer, data: torch.Tensor) -> None:
        self.initialize_parameters(quantizer, data)
        with torch.no_grad():
            tile = quantizer.granularity.tile_size(data.shape)
            # the running range may be wider than the data (an fp32 range continued with bf16 data):
            # the kernel reduces in the data dtype and merges into the range's dtype, as torch.min promotes
            ops.running_minmax_update_(self.min, self.max, data.detach(), tile, self.flags)
            if self._eager:
                self.check_finite()
        quantizer.quantization_range = (self.min, self.max)

    def check_finite(self) -> None:
        if self.flags is not None:
            _raise_for_flags(int(self.flags.item()))

    # ---- fused step: min/max + running update + range->params + int8 quantize in one kernel ----
    def _fused_mode(self, quantizer, callback, data) -> int:
        from ..nn.linear_quantizer import LinearQuantizer
        from ..quantized_tensor import QuantizedTensor

        if not self._fused or self._disable_quantization or type(quantizer) is not LinearQuantizer:
            return 0
        if quantizer.quantized_dtype != torch.int8 or quantizer.num_bits > 8 or torch.is_grad_enabled():
            return 0
        if not isinstance(data, torch.Tensor) or isinstance(data, QuantizedTensor) or not data.is_cuda:
            return 0
        if _flags.get_export_mode():
            return 0
        # the quantizer's own quantize must be what runs next (no other overrides in between)
        if not isinstance(callback, _Chain) or callback._pending or \
                getattr(callback._fn, "__func__", None) is not LinearQuantizer.quantize:
            return 0
        for p in (quantizer.scale, quantizer.offset):
            if p is not None and (p.dtype != torch.float32 or p.device != data.device):
                return 0
        return ops.calibrate_quantize_mode(data.shape, quantizer.granularity.tile_size(data.shape), data.dtype)

    def _fused_step(self, quantizer, data: torch.Tensor, mode: int):
        from ..quantization.affine.function import AffineQuantizationFunction
        from ..quantization.function import QuantizationContext
        from ..quantized_tensor import QuantizedTensor

        self.initialize_parameters(quantizer, data)
        if quantizer.has_uninitialized_params:
            quantizer._initialize_parameters(self.min.numel())
        one_sided_live = quantizer.symmetric and quantizer.allow_one_sided
        if mode == 1 and one_sided_live and self._settled is None:
            self._settled = torch.zeros(1, dtype=torch.int32, device=data.device)
            self._settled_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        if self._settled_host is not None and not self._settled_seen:
            # a plain host read of the pinned mirror: no sync; a stale 0 only costs one more fix-up launch
            self._settled_seen = bool(self._settled_host[0] != 0)
        tile = quantizer.granularity.tile_size(data.shape)
        want_rowsum = data.dim() >= 2 and (mode == 2 or (mode == 1 and self.min.numel() * data.shape[-1] == data.numel()))
        codes, rowsum = ops.calibrate_quantize_(
            self.min, self.max, data.detach(), tile, quantizer.num_bits, quantizer.symmetric, quantizer.allow_one_sided,
            quantizer.scale.data, None if quantizer.offset is None else quantizer.offset.data,
            self.flags, self._settled if mode == 1 else None, rowsum=want_rowsum, run_fixup=not self._settled_seen,
            workspace=self._arena.barrier_workspace(data.device) if (mode in (2, 3) and self._arena is not None) else None)
        if self._settled_host is not None and not self._settled_seen:
            self._settled_host.copy_(self._settled, non_blocking=True)
        if self._eager:
            self.check_finite()
        params = quantizer.quantization_parameters()
        params = params.with_changes(dequantize_dtype=params.dequantize_dtype or data.dtype)
        out = QuantizedTensor(codes, QuantizationContext(AffineQuantizationFunction, params))
        if rowsum is not None:
            out._ffq_rowsum = rowsum          # consumed by nn/qlinear.py instead of a separate row-sum pass
        return out

    def forward(self, quantizer, callback, args: tuple, kwargs: dict):
        data = args[0] if args else kwargs.get("data", next(iter(kwargs.values()), None))
        mode = self._fused_mode(quantizer, callback, data)
        if mode and self.min is not None and (self.min.device != data.device or
                                              torch.promote_types(self.min.dtype, data.dtype) != self.min.dtype):
            mode = 0     # a continued range on another device / in a narrower dtype: the plain path converts it
        if not mode:
            return super().forward(quantizer, callback, args, kwargs)
        if not self._initialized:
            self.setup_estimator(data)
            self._initialized = True
        return self._fused_step(quantizer, data, mode)

    def extra_repr(self) -> str:
        return f"min={self.min}, max={self.max}"


def _raise_for_flags(value: int) -> None:
    if value & 2:
        raise RuntimeError("fastforward_b200: the grid barrier of the fused calibration kernel timed out "
                           "(its workspace was shared between streams?)")
    if value & 1:
        raise NotImplementedError("Infinite")


class SmoothedMinMaxEstimator(SimpleEstimatorStep, torch.nn.Module):
    """EMA of the per-batch min/max: ``gamma*batch + (1-gamma)*running``, restarted whenever the
    running range still contains an infinity (minmax.py:79-90).  The per-tile reduction is the
    CUDA kernel; the EMA itself is a parameter-sized torch expression."""

    def __init__(self, quantizer, gamma: float = 1.0, disable_quantization: bool = False) -> None:
        super().__init__(disable_quantization=disable_quantization)
        lo, hi = quantizer.quantization_range
        self.gamma = gamma
        self.register_buffer("min", None if lo is None else lo.detach().clone())
        self.register_buffer("max", None if hi is None else hi.detach().clone())

    def estimate_step(self, quantizer, data: torch.Tensor) -> None:
        with torch.no_grad():
            tile = quantizer.granularity.tile_size(data.shape)
            batch_min, batch_max = ops.tile_minmax(data.detach(), tile)
            if self.min is None or self.max is None:
                self.min, self.max = batch_min, batch_max
            else:
                restart = torch.logical_or(self.min.isinf().any(), self.max.isinf().any())
                new_min = self.gamma * batch_min + (1 - self.gamma) * self.min
                new_max = self.gamma * batch_max + (1 - self.gamma) * self.max
                self.min = torch.where(restart, batch_min.to(new_min.dtype), new_min)
                self.max = torch.where(restart, batch_max.to(new_max.dtype), new_max)
        quantizer.quantization_range = (self.min, self.max)

    def extra_repr(self) -> str:
        return f"min={self.min}, max={self.max}"


class _MinMaxRangeEstimatorBase(RangeEstimator):
    skip_unsupported_quantizers = False

    def _check(self, module) -> None:
        if not isinstance(module, RangeSettable):
            proto = f"{RangeSettable.__module__}.{RangeSettable.__qualname__}"
            raise TypeError(f"{type(module).__name__} does not implement {proto}.")

    def cleanup(self, module, metadata) -> None:
        del module
        metadata.remove()

    def split_module(self, module: torch.nn.Module) -> Iterator[Quantizer]:
        quantizers = [("", module)] if isinstance(module, Quantizer) and not module.is_stub() else []
        quantizers += [(n, q) for n, q in named_quantizers(module, recurse=True) if q is not module]
        for _, quantizer in quantizers:
            if isinstance(quantizer, RangeSettable) or not self.skip_unsupported_quantizers:
                yield quantizer
            else:
                logger.warning(f"{type(quantizer).__name__} does not implement RangeSettable. Therefore it is not "
                               f"included in {type(self).__name__} range setting.")


class RunningMinMaxRangeEstimator(_MinMaxRangeEstimatorBase):
    def __init__(self, disable_quantization: bool = False, skip_unsupported_quantizers: bool = False, *,
                 eager_checks: bool = False, sync_ranges: bool = False, process_group=None, fused: bool = True) -> None:
        self.disable_quantization = disable_quantization
        self.skip_unsupported_quantizers = skip_unsupported_quantizers
        self.eager_checks = eager_checks
        self.sync_ranges = sync_ranges or process_group is not None
        self.process_group = process_group
        self.fused = fused
        self._steps: list = []
        self._arena = _RangeArena()

    def prepare(self, module):
        self._check(module)
        step = RunningMinMaxEstimator(module, disable_quantization=self.disable_quantization,
                                      eager_checks=self.eager_checks, arena=self._arena, fused=self.fused)
        self._steps.append((module, step))
        return module.register_override(step)

    def finalize(self, prepared: Sequence[tuple]) -> None:
        del prepared
        steps = [(q, s) for q, s in self._steps if s.min is not None]
        flags = list({id(s.flags): s.flags for _, s in steps if s.flags is not None}.values())
        if self.sync_ranges:
            from ..distributed import all_reduce_minmax_buffers, all_reduce_ranges

            arena_ids = {id(mn.untyped_storage()) for mn, _ in self._arena.buffers()}
            loose = [(s.min, s.max) for _, s in steps if id(s.min.untyped_storage()) not in arena_ids]
            all_reduce_minmax_buffers(list(self._arena.buffers()), flags, group=self.process_group)
            all_reduce_ranges(loose, None, group=self.process_group)
            self._reset_parameters_from_ranges(steps)
        # ONE host sync for all quantizers: the reference's per-step `isinf().any()` checks
        if flags and not self.eager_checks:
            value = 0
            for v in torch.stack([f.reshape(()) for f in flags]).tolist():
                value |= int(v)
            _raise_for_flags(value)
        self._steps = []
        self._arena = _RangeArena()


    def _reset_parameters_from_ranges(self, steps) -> None:
        """After the ranges were merged across ranks: every quantizer's (scale, offset) from its merged range.  The
        quantizers whose running range lives in an arena chunk and whose parameters are materialised fp32 tensors of
        the right size are done with ONE launch per chunk; the rest go through the ``quantization_range`` setter."""
        from ..nn.linear_quantizer import LinearQuantizer

        chunks = {}
        for chunk_list in self._arena.chunks.values():
            for mn, mx, used in chunk_list:
                if used:
                    chunks[mn.untyped_storage().data_ptr()] = (mn, mx, [])
        for quantizer, step in steps:
            entry = chunks.get(step.min.untyped_storage().data_ptr())
            n = step.min.numel()
            ok = entry is not None and type(quantizer) is LinearQuantizer and not quantizer.has_uninitialized_params \
                and step.min.is_contiguous() and step.max.storage_offset() == step.min.storage_offset() \
                and quantizer.scale.dtype == torch.float32 and quantizer.scale.numel() == n \
                and quantizer.scale.device == step.min.device and quantizer.scale.is_contiguous() \
                and (quantizer.offset is None or (quantizer.offset.dtype == torch.float32 and quantizer.offset.numel() == n
                                                  and quantizer.offset.is_contiguous()))
            if not ok:
                quantizer.quantization_range = (step.min, step.max)
                continue
            entry[2].append((step.min.storage_offset(), n, quantizer.num_bits, quantizer.symmetric, quantizer.allow_one_sided,
                             quantizer.scale.data, None if quantizer.offset is None else quantizer.offset.data))
        for mn, mx, entries in chunks.values():
            ops.parameters_for_ranges_batched_(mn, mx, entries)


class SmoothedMinMaxRangeEstimator(_MinMaxRangeEstimatorBase):
    def __init__(self, gamma: float = 1.0, disable_quantization: bool = False,
                 skip_unsupported_quantizers: bool = False) -> None:
        self.gamma = gamma
        self.disable_quantization = disable_quantization
        self.skip_unsupported_quantizers = skip_unsupported_quantizers

    def prepare(self, module):
        self._check(module)
        return module.register_override(
            SmoothedMinMaxEstimator(module, gamma=self.gamma, disable_quantization=self.disable_quantization))


running_minmax = RunningMinMaxRangeEstimator
smoothed_minmax = SmoothedMinMaxRangeEstimator
