"""Fuse quantize-dequantize ("QDQ", grid-snapped) values into a model's weights, in place
(reference: quantization/fuse.py:91-121,199-241).

The reference snaps each weight with ``weight.copy_(quantizer(weight).dequantize())`` -- a quantize
kernel chain, a dequantize chain and a copy, i.e. >= 5 passes over the weight.  Here a
``LinearQuantizer`` takes the fused fake-quantize kernel writing straight into the weight's own
storage: one read and one write per element (2s bytes), bit-identical to the two-step result.
``rank``/``world_size`` shard the independent targets by ``i mod world`` (BASELINE config 3:
whole-model W4 g=128 weight quantization sharded by layer across GPUs, no collective)."""

from __future__ import annotations

import ctypes
import weakref
from typing import Iterator, List, Optional, Protocol, Tuple, runtime_checkable

import torch

from .. import _cabi as C
from .. import flags
from ..exceptions import QuantizationError
from ..nn.linear_quantizer import LinearQuantizer
from ..nn.quantizer import Quantizer, QuantizerStub
from ..quant_init import find_quantizers

WeightQuantizerTarget = Tuple[torch.nn.Module, str, Quantizer]


@runtime_checkable
class WeightQuantizerDiscovery(Protocol):
    """Anything callable as ``discovery(model) -> iterator of (module, weight_attr, quantizer)`` (fuse.py:36-50)."""

    def __call__(self, model: torch.nn.Module) -> Iterator[WeightQuantizerTarget]:
        ...


class ConventionDiscovery:
    """``(module, weight_attr, quantizer)`` for every initialised quantizer tagged ``tag`` whose
    parent has a ``weight_attr`` Parameter."""

    def __init__(self, weight_attr: str = "weight", *, tag: str = "parameter/weight") -> None:
        self._weight_attr, self._tag = weight_attr, tag

    def __call__(self, model: torch.nn.Module) -> Iterator[WeightQuantizerTarget]:
        for result in find_quantizers(model, f"**/[quantizer:{self._tag}]"):
            quantizer = result.module
            if quantizer.is_stub():
                continue
            weight = getattr(result.parent, self._weight_attr, None)
            if isinstance(weight, torch.nn.Parameter):
                yield result.parent, self._weight_attr, quantizer


def _check_tied(targets: List[WeightQuantizerTarget]) -> None:
    by_weight = {}
    for module, attr, quantizer in targets:
        by_weight.setdefault(id(getattr(module, attr)), []).append(quantizer)
    for quantizers in by_weight.values():
        if len({id(q) for q in quantizers}) > 1:
            raise QuantizationError(
                "Cannot fuse QDQ weights: a weight is tied across modules whose weight quantizers snap it to "
                "different grids. Untie the weights or share a single quantizer instance between them.")


def _fused_inplace(weight: torch.nn.Parameter, q: LinearQuantizer) -> bool:
    """weight <- dequantize(quantize(weight)) with ONE kernel writing into the weight's storage."""
    w = weight.data
    if not (w.is_cuda and w.is_contiguous() and w.dtype.is_floating_point) or q.has_uninitialized_params:
        return False
    tile = q.granularity.tile_size(w.shape)
    tile = tuple(w.shape) if isinstance(tile, str) else tuple(tile)
    layout = C.make_layout(tuple(w.shape), tile)
    scale = q.scale.detach().reshape(-1).contiguous()
    offset = None if q.offset is None else q.offset.detach().reshape(-1).contiguous()
    qdtype = q.quantized_dtype or w.dtype
    C.check(C.lib.ffq_fakequant_fwd(
        w.data_ptr(), C.dtype_tag(w.dtype), w.data_ptr(), C.dtype_tag(w.dtype), None, C.dtype_tag(qdtype),
        scale.data_ptr(), C.dtype_tag(scale.dtype), C.ptr(offset),
        C.dtype_tag(offset.dtype if offset is not None else None),
        ctypes.byref(layout), float(q.num_bits), C.current_stream(w.device)))
    return True


def _fuse_target(module: torch.nn.Module, weight_attr: str, quantizer: Quantizer, *, stub_quantizer: bool) -> None:
    weight = getattr(module, weight_attr)
    done = False
    if type(quantizer) is LinearQuantizer and not list(quantizer.overrides):
        done = _fused_inplace(weight, quantizer)     # elementwise, so in-place is safe
    if not done:
        with flags.strict_quantization(False):
            qdq = quantizer(weight)
            qdq = qdq.dequantize() if hasattr(qdq, "quant_args") else qdq
        if qdq is not weight:
            with torch.no_grad():
                weight.copy_(qdq)
    if stub_quantizer:
        for name, child in list(module.named_children()):
            if child is quantizer:
                setattr(module, name, QuantizerStub(_metadata=quantizer.quant_metadata))


def find_weight_quantizers(model: torch.nn.Module, *, discovery=None) -> List[WeightQuantizerTarget]:
    """The ``(module, weight_attr, quantizer)`` targets a fuse or stub pass would act on, without performing it
    (fuse.py:244-264)."""
    discovery = ConventionDiscovery() if discovery is None else discovery
    return list(discovery(model))


def stub_weight_quantizers(model: torch.nn.Module, *, discovery=None) -> None:
    """Replace the discovered weight quantizers by stubs carrying their metadata; weights are neither read nor written
    (fuse.py:267-300).  For models whose saved weights are already grid-snapped: the forward pass must not quantize
    them a second time."""
    for module, _, quantizer in find_weight_quantizers(model, discovery=discovery):
        for name, child in list(module.named_children()):
            if child is quantizer:
                setattr(module, name, QuantizerStub(_metadata=quantizer.quant_metadata))
                break


def fuse_qdq_weights(model: torch.nn.Module, *, stub_quantizers: bool = False, discovery=None,
                     rank: Optional[int] = None, world_size: Optional[int] = None) -> int:
    """Snap every discovered weight to its quantization grid in place; returns how many were fused
    by this rank.  With ``world_size`` > 1 only targets ``i % world_size == rank`` are processed."""
    discovery = ConventionDiscovery() if discovery is None else discovery
    targets = list(discovery(model))
    _check_tied(targets)
    if world_size is not None and world_size > 1:
        from ..distributed import shard_units

        targets = [targets[i] for i in shard_units(len(targets), rank, world_size)]
    for module, attr, quantizer in targets:
        _fuse_target(module, attr, quantizer, stub_quantizer=stub_quantizers)
    return len(targets)


def calibrate_weight_quantizers(model: torch.nn.Module, discovery=None, rank: Optional[int] = None,
                                world_size: Optional[int] = None) -> int:
    """Set each weight quantizer's range from its own weight (min/max kernel + params kernel):
    the weight half of a calibration pass, without running the model."""
    from .. import ops

    discovery = ConventionDiscovery() if discovery is None else discovery
    targets = list(_all_targets(model, discovery))
    if world_size is not None and world_size > 1:
        from ..distributed import shard_units

        targets = [targets[i] for i in shard_units(len(targets), rank, world_size)]
    for module, attr, quantizer in targets:
        w = getattr(module, attr).detach()
        tile = quantizer.granularity.tile_size(w.shape)
        lo, hi = ops.tile_minmax(w, tile)
        quantizer.quantization_range = (lo, hi)
    return len(targets)


# model -> {configuration: FakeQuantBatch} of its last whole-model call, reused while the storages stay where they are
_BATCH_CACHE: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def calibrate_and_fuse_qdq_weights(model: torch.nn.Module, *, stub_quantizers: bool = False, discovery=None,
                                   rank: Optional[int] = None, world_size: Optional[int] = None, batched: bool = True) -> int:
    """``calibrate_weight_quantizers`` followed by ``fuse_qdq_weights`` -- the unit of whole-model weight quantization
    -- with the weight read ONCE: per-tile min/max -> the quantizer's scale/offset -> the weight snapped to its grid in
    place (``ffq_calibrate_fakequant``) where the layout allows it (per-channel rows, per-group tiles), the two
    separate steps otherwise.  Per-group weights of one dtype and one quantizer configuration (the usual W4 g=128
    recipe over every linear of a model) go through ONE launch for all of them (``ffq_calibrate_fakequant_batched``;
    ``batched=False`` keeps one launch per weight).  Bit-identical to the two-step sequence."""
    from .. import ops

    discovery = ConventionDiscovery() if discovery is None else discovery
    targets = list(_all_targets(model, discovery))
    _check_tied(targets)
    if world_size is not None and world_size > 1:
        from ..distributed import shard_units

        targets = [targets[i] for i in shard_units(len(targets), rank, world_size)]
    # ---- per-group 16-bit weights that share a configuration: one launch for all of them ------------------------
    done = set()
    if batched:
        groups: dict = {}
        seen_weights = set()
        for idx, (module, attr, quantizer) in enumerate(targets):
            w = getattr(module, attr).data
            if type(quantizer) is not LinearQuantizer or list(quantizer.overrides) or not w.is_cuda or not w.is_contiguous() \
                    or w.dtype not in (torch.bfloat16, torch.float16) or w.dim() != 2 or id(getattr(module, attr)) in seen_weights:
                continue
            tile = quantizer.granularity.tile_size(w.shape)
            if isinstance(tile, str) or tuple(tile[:-1]) != (1,) * (w.dim() - 1) or tile[-1] not in (64, 128) or w.data_ptr() % 32:
                continue
            n = quantizer.granularity.parameter_dimensionality(w.shape)
            if quantizer.has_uninitialized_params:
                quantizer._initialize_parameters(n)
            params = [quantizer.scale] + ([] if quantizer.offset is None else [quantizer.offset])
            if not all(p.dtype == torch.float32 and p.device == w.device and p.numel() == n and p.is_contiguous() for p in params):
                continue
            seen_weights.add(id(getattr(module, attr)))
            key = (w.device, w.dtype, tile[-1], quantizer.num_bits, quantizer.symmetric, quantizer.allow_one_sided,
                   quantizer.quantized_dtype)
            groups.setdefault(key, []).append(idx)
        cache = _BATCH_CACHE.setdefault(model, {})
        for key, idxs in groups.items():
            if len(idxs) < 2:
                continue
            ws = [getattr(targets[i][0], targets[i][1]).data for i in idxs]
            ss = [targets[i][2].scale.data for i in idxs]
            os_ = [None if targets[i][2].offset is None else targets[i][2].offset.data for i in idxs]
            sig = tuple((w.data_ptr(), s.data_ptr(), 0 if o is None else o.data_ptr()) for w, s, o in zip(ws, ss, os_))
            batch = cache.get(key)
            if batch is None or batch.sig != sig:
                try:
                    batch = ops.FakeQuantBatch(ws, ss, os_, key[2])
                except NotImplementedError:
                    continue
                batch.sig = sig
                cache[key] = batch
            try:
                with torch.no_grad():
                    ops.calibrate_fake_quantize_batched_(batch, key[3], key[4], key[5], key[6])
            except NotImplementedError:      # e.g. a code dtype that rounds the codes: the per-tensor path decides
                continue
            done.update(idxs)
    for idx, (module, attr, quantizer) in enumerate(targets):
        if idx in done:
            continue
        weight = getattr(module, attr)
        w = weight.data
        tile = quantizer.granularity.tile_size(w.shape)
        fused = type(quantizer) is LinearQuantizer and not list(quantizer.overrides) and w.is_cuda and w.is_contiguous() \
            and ops.calibrate_quantize_mode(w.shape, tile, w.dtype) in (1, 3)
        if fused:
            n = quantizer.granularity.parameter_dimensionality(w.shape)
            if quantizer.has_uninitialized_params:
                quantizer._initialize_parameters(n)
            params = [quantizer.scale] + ([] if quantizer.offset is None else [quantizer.offset])
            fused = all(p.dtype == torch.float32 and p.device == w.device and p.numel() == n and p.is_contiguous() for p in params)
        if fused:
            try:
                with torch.no_grad():
                    ops.calibrate_fake_quantize_(w, tile, quantizer.num_bits, quantizer.symmetric, quantizer.allow_one_sided,
                                                 quantizer.scale.data, None if quantizer.offset is None else quantizer.offset.data,
                                                 quantizer.quantized_dtype, out=w)
            except NotImplementedError:          # e.g. a code dtype that rounds the codes: keep the exact two-step path
                fused = False
        if not fused:
            lo, hi = ops.tile_minmax(w.detach(), tile)
            quantizer.quantization_range = (lo, hi)
            _fuse_target(module, attr, quantizer, stub_quantizer=False)
    if stub_quantizers:
        for module, attr, quantizer in targets:
            for name, child in list(module.named_children()):
                if child is quantizer:
                    setattr(module, name, QuantizerStub(_metadata=quantizer.quant_metadata))
    return len(targets)


def _all_targets(model: torch.nn.Module, discovery) -> Iterator[WeightQuantizerTarget]:
    """Like the discovery, but includes quantizers whose parameters are still uninitialised."""
    tag = getattr(discovery, "_tag", "parameter/weight")
    attr = getattr(discovery, "_weight_attr", "weight")
    for result in find_quantizers(model, f"**/[quantizer:{tag}]"):
        if result.module.is_stub():
            continue
        if isinstance(getattr(result.parent, attr, None), torch.nn.Parameter):
            yield result.parent, attr, result.module

