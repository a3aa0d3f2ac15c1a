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
    forward) and the layout is a per-channel row, a whole tensor or a per-group tile, estimate_step +
    range setter + quantize (+ the code row sums the W8A8 linear needs) are ONE kernel reading the tensor
    once (``ops.calibrate_quantize_``; pass ``fused=False`` to the estimator to keep the separate kernels);
  * quantizers that are handed the SAME tensor object with the same configuration (the q/k/v and the
    gate/up input quantizers of a decoder layer) share one launch and one set of codes (``dedupe``);
  * a quantizer whose input is an unchanged ``nn.Parameter`` (same storage, same version counter) since
    its previous step returns its previous codes: the running range cannot move, so parameters and
    codes are bit-identical (``memoize_parameters``);
  * optionally (``overlap_parameters=N``) the fused step of a weight is launched N weight quantizers ahead of its
    use on a side stream -- a parallel branch of a captured graph; measured: no gain on a power-capped B200;
  * when the block ends, everything the host can do without the device's answer (the slot signature and its
    all-reduce, the descriptor table of the batched parameters launch, storage for de-duplicated quantizers, the
    removal of the overrides) happens BEFORE the exit's single host sync, while the device still works through the
    queued steps; after the sync: two all-reduces, one launch, one multi-tensor copy.

The estimator recognises the unmodified reference's ``LinearQuantizer`` and override chain as well as this
package's, so ``plugin.install(patch_estimators=True)`` puts all of this under ``fastforward.estimate_ranges``."""

from __future__ import annotations

import logging
import sys
import types
import weakref
from typing import Iterator, List, Optional, Sequence

import torch

from .. import ops
from ..nn.quantizer import Quantizer
from .common import RangeEstimator, RangeSettable, SimpleEstimatorStep

logger = logging.getLogger(__name__)

_OWN_ROOT = __name__.partition(".")[0]


def _param_count(quantizer, data: torch.Tensor) -> int:
    return quantizer.granularity.parameter_dimensionality(data.shape)


# ---------------------------------------------------------------------------------------------------------------
# host-package adapters: this package's classes, or an installed unmodified ``fastforward``
# ---------------------------------------------------------------------------------------------------------------
_HOSTS: dict = {}


def _host_of(quantizer):
    """Classes of the package a quantizer belongs to (QuantizedTensor, QuantizationContext, export flag)."""
    root = type(quantizer).__module__.partition(".")[0]
    host = _HOSTS.get(root)
    if host is None:
        mod = sys.modules.get(root)
        fn_mod = sys.modules.get(f"{root}.quantization.function")
        lq_mod = sys.modules.get(f"{root}.nn.linear_quantizer")
        if mod is None or fn_mod is None or not hasattr(mod, "QuantizedTensor"):
            # a quantizer from neither host package (a user's own RangeSettable): separate kernels + its own setter
            host = types.SimpleNamespace(QuantizedTensor=(), QuantizationContext=None, get_export_mode=lambda: False,
                                         LinearQuantizer=None, rcp=False)
        else:
            host = types.SimpleNamespace(
                QuantizedTensor=mod.QuantizedTensor, QuantizationContext=fn_mod.QuantizationContext,
                get_export_mode=mod.get_export_mode, LinearQuantizer=getattr(lq_mod, "LinearQuantizer", None),
                # an unmodified reference computes parameters_for_range with aten's CUDA kernels on a GPU (its scalar
                # divisions multiply by the reciprocal): its quantizers get that flavour, ours the CPU flavour the
                # golden vectors pin
                rcp=root != _OWN_ROOT)
        _HOSTS[root] = host
    return host


def _is_quantizer(module) -> bool:
    if isinstance(module, Quantizer):
        return True
    return isinstance(module, torch.nn.Module) and all(
        callable(getattr(module, name, None)) for name in ("register_override", "is_stub", "quantize"))


def _next_is_own_quantize(callback, quantizer) -> bool:
    """True when calling ``callback`` would run ``quantizer.quantize`` and nothing else (no other override in
    between): this package's ``_Chain`` or the reference's ``_WrappedOverriddenFn`` (forward_override.py:76-96)."""
    pending = getattr(callback, "_pending", None)
    fn = getattr(callback, "_fn", None)
    if pending is None:
        pending = getattr(callback, "override_stack", None)
        fn = getattr(callback, "overridden_fn", None)
    if pending is None or pending:
        return False
    return getattr(fn, "__self__", None) is quantizer and getattr(fn, "__func__", None) is type(quantizer).quantize


def _set_range_direct(quantizer, lo: torch.Tensor, hi: torch.Tensor) -> bool:
    """``quantizer.quantization_range = (lo, hi)`` for a LinearQuantizer of either host with ONE sync-free kernel
    writing scale/offset in place (nn/linear_quantizer.py:327-357).  False: the caller uses the property setter."""
    host = _host_of(quantizer)
    if host.LinearQuantizer is None or type(quantizer) is not host.LinearQuantizer or not lo.is_cuda:
        return False
    if quantizer.has_uninitialized_params:
        quantizer._initialize_parameters(lo.numel())
    scale, offset = quantizer.scale, quantizer.offset
    n = lo.numel()
    for p in (scale, offset):
        if p is not None and (p.dtype != torch.float32 or p.device != lo.device or p.numel() != n or not p.is_contiguous()):
            return False
    with torch.no_grad():
        ops.parameters_for_range_(lo, hi, quantizer.num_bits, quantizer.symmetric, quantizer.allow_one_sided,
                                  scale.data, None if offset is None else offset.data, reciprocal_scalar_division=host.rcp)
    return True


# ---------------------------------------------------------------------------------------------------------------
# per-block state shared by all steps of one estimate_ranges block
# ---------------------------------------------------------------------------------------------------------------
class _RangeArena:
    """Running ranges of many quantizers live in a few flat buffers (one min and one max buffer per dtype/device), so
    the data-parallel exchange at the end of the block is two collectives on contiguous memory -- no packing of
    hundreds of tiny tensors -- and one flag word serves every quantizer."""

    CHUNK = 1 << 21

    def __init__(self, chunk: Optional[int] = None) -> None:
        self.chunk = chunk or self.CHUNK
        self.chunks: dict = {}     # (device, dtype) -> list of [min_buf, max_buf, used]
        self.flags: dict = {}      # device -> int32[1]
        self.barriers: dict = {}   # device -> zeroed scratch of the fused per-tensor kernel's grid barrier

    def take(self, n: int, like: torch.Tensor):
        """(min view, max view, (chunk index, offset)) of a fresh slot of n elements."""
        key = (like.device, like.dtype)
        chunks = self.chunks.setdefault(key, [])
        if not chunks or chunks[-1][2] + n > chunks[-1][0].numel():
            size = max(self.chunk, n)
            chunks.append([like.new_full((size,), float("inf")), like.new_full((size,), float("-inf")), 0])
        mn, mx, used = chunks[-1]
        chunks[-1][2] = used + n
        return mn[used:used + n], mx[used:used + n], (len(chunks) - 1, used)

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


class _BlockState:
    """What the steps of one block share: the arenas (activations apart from parameters: only the former differ
    between data-parallel ranks), the most recent fused outputs by input tensor (dedupe) and the counters."""

    def __init__(self, dedupe: bool, memoize: bool, overlap: int = 0) -> None:
        self.act = _RangeArena(chunk=1 << 14)
        self.param = _RangeArena()
        self.recent: dict = {}        # (data_ptr, shape, dtype) -> _Recent
        self.aliases: List["RunningMinMaxEstimator"] = []
        self.dedupe, self.memoize = dedupe, memoize
        self.stats = {"fused": 0, "deduped": 0, "memoized": 0, "separate": 0, "overlapped": 0}
        # overlapped parameter steps: the fused step of a weight does not depend on the activations, so it is launched
        # ``overlap`` parameter quantizers AHEAD of its use on a side stream (a parallel branch of a captured graph):
        # an HBM-bound kernel under the tensor-bound GEMMs of the layers before it
        self.overlap = 0 if memoize else int(overlap)
        self.porder: List["RunningMinMaxEstimator"] = []     # parameter steps in the order they first ran
        self.side: dict = {}                                  # device -> side stream

    def side_stream(self, device: torch.device):
        if device not in self.side:
            self.side[device] = torch.cuda.Stream(device=device)
        return self.side[device]

    def side_workspace(self, device: torch.device) -> torch.Tensor:
        key = ("ws", device)
        if key not in self.side:
            self.side[key] = torch.zeros(ops._CALQ_WS, dtype=torch.uint8, device=device)
        return self.side[key]

    def launch_ahead(self, step: "RunningMinMaxEstimator") -> None:
        """After ``step`` has been served: make sure the next ``overlap`` parameter steps are in flight."""
        order = self.porder
        for j in range(step._pidx + 1, min(len(order), step._pidx + 1 + self.overlap)):
            if order[j]._ahead is None:
                order[j]._issue_ahead()

    def drain(self) -> None:
        """Join every parameter step still in flight into the current stream (block exit, failed block)."""
        for s in self.porder:
            s._drop_ahead()

    def flag(self, device):
        return self.param.flag(device)        # one flag word per device for the whole block

    def all_flags(self):
        return list(self.param.flags.values())


class _Recent:
    __slots__ = ("ref", "version", "step", "codes", "rowsum")

    def __init__(self, data, step, codes, rowsum) -> None:
        self.ref, self.version, self.step, self.codes, self.rowsum = weakref.ref(data), data._version, step, codes, rowsum


def _same_config(a, b) -> bool:
    return (type(a) is type(b) and a.num_bits == b.num_bits and a.symmetric == b.symmetric
            and a.allow_one_sided == b.allow_one_sided and a.quantized_dtype == b.quantized_dtype
            and repr(a.granularity) == repr(b.granularity))


class RunningMinMaxEstimator(SimpleEstimatorStep, torch.nn.Module):
    def __init__(self, quantizer, disable_quantization: bool = False, eager_checks: bool = False,
                 arena: Optional[_RangeArena] = None, fused: bool = True, state: Optional[_BlockState] = None) -> None:
        super().__init__(disable_quantization=disable_quantization)
        self._state = state
        self._arena = arena               # explicit arena (tests); otherwise chosen from the state by the data's kind
        self._fused = fused
        self._quantizer_ref = weakref.ref(quantizer)
        self._settled = None          # device int32[1]: set by the kernel once every running min is negative
        self._settled_host = None     # pinned mirror, refreshed asynchronously: lets later steps skip the fix-up launch
        self._settled_seen = False
        self._nsteps = 0
        self._alias_of: Optional["RunningMinMaxEstimator"] = None
        self._memo = None             # (weakref(data), version, output) of the previous fused step on a Parameter
        self._ahead = None            # (event, weakref(data), version, output) of a step launched ahead on the side stream
        self._pidx = None             # position among the block's parameter steps (``_BlockState.porder``)
        self._pdata = None            # (weakref(data), version, mode) of this step's most recent launch on a Parameter
        self._slot = None             # (arena, chunk index, offset, n) of the running range when it lives in an arena
        self._param_data = False      # the data this quantizer sees is an nn.Parameter (identical on every DP rank)
        lo, hi = quantizer.quantization_range       # continues from an existing range (minmax.py:198-200)
        self._fresh = lo is None and hi is None
        self.register_buffer("min", None if lo is None else lo.detach().clone())
        self.register_buffer("max", None if hi is None else hi.detach().clone())
        self.register_buffer("flags", None)
        self._eager = eager_checks

    # ---- running range storage ---------------------------------------------------------------------------
    def _arena_for(self, data) -> Optional[_RangeArena]:
        if self._arena is not None:
            return self._arena
        if self._state is None:
            return None
        return self._state.param if self._param_data else self._state.act

    def initialize_parameters(self, quantizer, data: torch.Tensor) -> None:
        if self.min is not None and self.max is not None and self.flags is not None and \
                self.min.dtype == data.dtype and self.min.device == data.device:
            return                                  # steady state: nothing to do
        self._param_data = isinstance(data, torch.nn.Parameter)
        n = _param_count(quantizer, data)
        arena = self._arena_for(data)
        if self.min is None and self.max is None and arena is not None:
            self.min, self.max, (ci, off) = arena.take(n, data)
            self._slot = (arena, ci, off, n)
        if self.min is None:
            self.min = data.new_full((n,), float("inf"))
        if self.max is None:
            self.max = data.new_full((n,), float("-inf"))
        if self.min.dtype != data.dtype or self.min.device != data.device:
            # torch.min(self.min, data_min) in the reference promotes; keep the running range in the
            # promoted dtype so that nothing is lost
            dt = torch.promote_types(self.min.dtype, data.dtype)
            if dt != self.min.dtype or self.min.device != data.device:
                self.min, self.max = self.min.to(device=data.device, dtype=dt), self.max.to(device=data.device, dtype=dt)
                self._slot = None
        if self.flags is None:
            if self._state is not None:
                self.flags = self._state.flag(data.device)
            elif self._arena is not None:
                self.flags = self._arena.flag(data.device)
            else:
                self.flags = torch.zeros(1, dtype=torch.int32, device=data.device)

    def estimate_step(self, quantizer, data: torch.Tensor) -> None:
        self._unalias()
        self.initialize_parameters(quantizer, data)
        with torch.no_grad():
            tile = quantizer.granularity.tile_size(data.shape)
            # the running range may be wider than the data (an fp32 range continued with bf16 data):
            # the kernel reduces in the data dtype and merges into the range's dtype, as torch.min promotes
            ops.running_minmax_update_(self.min, self.max, data.detach(), tile, self.flags)
            if self._eager:
                self.check_finite()
        self._nsteps += 1
        if self._state is not None:
            self._state.stats["separate"] += 1
        if not _set_range_direct(quantizer, self.min, self.max):
            quantizer.quantization_range = (self.min, self.max)

    def check_finite(self) -> None:
        if self.flags is not None:
            _raise_for_flags(int(self.flags.item()))

    # ---- fused step: min/max + running update + range->params + int8 quantize in one kernel ----
    def _fused_mode(self, quantizer, callback, data) -> int:
        if not self._fused or self._disable_quantization:
            return 0
        host = _host_of(quantizer)
        if host.LinearQuantizer is None or type(quantizer) is not host.LinearQuantizer:
            return 0
        if quantizer.quantized_dtype != torch.int8 or quantizer.num_bits > 8 or torch.is_grad_enabled():
            return 0
        if not isinstance(data, torch.Tensor) or isinstance(data, host.QuantizedTensor) or not data.is_cuda:
            return 0
        if host.get_export_mode():
            return 0
        # the quantizer's own quantize must be what runs next (no other overrides in between)
        if not _next_is_own_quantize(callback, quantizer):
            return 0
        for p in (quantizer.scale, quantizer.offset):
            if p is not None and (p.dtype != torch.float32 or p.device != data.device):
                return 0
        return ops.calibrate_quantize_mode(data.shape, quantizer.granularity.tile_size(data.shape), data.dtype)

    def _wrap(self, quantizer, codes, rowsum, data_dtype):
        host = _host_of(quantizer)
        params = quantizer.quantization_parameters()
        params = params.with_changes(dequantize_dtype=params.dequantize_dtype or data_dtype)
        out = host.QuantizedTensor(codes, host.QuantizationContext(quantizer.quantization_function, params))
        if rowsum is not None:
            out._ffq_rowsum = rowsum          # consumed by nn/qlinear.py instead of a separate row-sum pass
        return out

    def _become_alias(self, quantizer, master: "RunningMinMaxEstimator") -> None:
        """Share the master's running range and (for the length of the block) its parameter storage: both quantizers
        have seen exactly the same tensors, so everything derived from them is equal by construction."""
        mq = master._quantizer_ref()
        n = master.min.numel()
        if quantizer.has_uninitialized_params:
            quantizer._initialize_parameters(n)
        self.min, self.max, self.flags = master.min, master.max, master.flags
        self._slot, self._param_data = None, master._param_data
        with torch.no_grad():
            quantizer.scale.data = mq.scale.data
            if quantizer.offset is not None:
                quantizer.offset.data = mq.offset.data
        self._alias_of = master
        self._state.aliases.append(self)

    def _unalias(self) -> None:
        """Give an aliased quantizer its own copies back (values unchanged)."""
        if self._alias_of is None:
            return
        q = self._quantizer_ref()
        self._alias_of = None
        self.min, self.max = self.min.clone(), self.max.clone()
        if q is not None:
            with torch.no_grad():
                q.scale.data = q.scale.data.clone()
                if q.offset is not None:
                    q.offset.data = q.offset.data.clone()

    def _fused_step(self, quantizer, data: torch.Tensor, mode: int):
        st = self._state
        # (1) an unchanged parameter since this quantizer's previous step: same range, same parameters, same codes
        if self._memo is not None:
            ref, version, out = self._memo
            if ref() is data and data._version == version:
                self._nsteps += 1
                st.stats["memoized"] += 1
                return out
            self._memo = None
        # (1b) this step was launched ahead on the side stream (overlapped parameter steps)
        if self._ahead is not None:
            result = self._take_ahead(data)
            if result is not None:
                codes, rowsum = result
                self._nsteps += 1
                st.stats["overlapped"] += 1
                if st.dedupe:
                    st.recent[(data.data_ptr(), data.shape, data.dtype)] = _Recent(data, self, codes, rowsum)
                st.launch_ahead(self)
                return self._wrap(quantizer, codes, rowsum, data.dtype)
        # (2) another quantizer of the same configuration has just processed this very tensor
        key = None
        if st is not None and st.dedupe:
            key = (data.data_ptr(), data.shape, data.dtype)
            ent = st.recent.get(key)
            if ent is not None and ent.step is not self and ent.ref() is data and ent.version == data._version:
                master = ent.step
                hit = self._alias_of is master and master._nsteps == self._nsteps + 1
                if not hit and self._alias_of is None and self._nsteps == 0 and master._nsteps == 1 and self._fresh \
                        and master._fresh and master._alias_of is None and self.min is None \
                        and _same_config(quantizer, master._quantizer_ref()):
                    self._become_alias(quantizer, master)
                    hit = True
                if hit:
                    self._nsteps += 1
                    st.stats["deduped"] += 1
                    return self._wrap(quantizer, ent.codes, ent.rowsum, data.dtype)
        codes, rowsum = self._launch(quantizer, data, mode)
        out = self._wrap(quantizer, codes, rowsum, data.dtype)
        if st is not None:
            st.stats["fused"] += 1
            if key is not None:
                st.recent[key] = _Recent(data, self, codes, rowsum)
            if st.memoize and isinstance(data, torch.nn.Parameter):
                self._memo = (weakref.ref(data), data._version, out)
            if st.overlap and mode in (1, 3) and isinstance(data, torch.nn.Parameter):
                if self._pidx is None:
                    self._pidx = len(st.porder)
                    st.porder.append(self)
                self._pdata = (weakref.ref(data), data._version, mode)
                st.launch_ahead(self)
        return out

    def _launch(self, quantizer, data: torch.Tensor, mode: int, stream=None):
        """The fused kernel itself (on ``stream`` when given, else on the current stream) -> (codes, rowsum)."""
        self._unalias()
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
        # row sums ride along when the rows are whole 16-byte vectors (the kernels' vector width); other shapes get
        # them from ffq_rowsum_i8 inside the linear
        vec = 16 // data.element_size()
        want_rowsum = data.dim() >= 2 and data.shape[-1] % vec == 0 and \
            (mode == 2 or (mode == 1 and self.min.numel() * data.shape[-1] == data.numel()))
        ws = None
        if mode in (2, 3):
            arena = self._arena_for(data)
            ws = arena.barrier_workspace(data.device) if arena is not None else None
            if stream is not None:            # the side stream's launches keep their flag words apart from the main stream's
                ws = self._state.side_workspace(data.device)
        codes, rowsum = ops.calibrate_quantize_(
            self.min, self.max, data.detach(), tile, quantizer.num_bits, quantizer.symmetric, quantizer.allow_one_sided,
            quantizer.scale.data, None if quantizer.offset is None else quantizer.offset.data,
            self.flags, self._settled if mode == 1 else None, rowsum=want_rowsum, run_fixup=not self._settled_seen,
            workspace=ws, reciprocal_scalar_division=_host_of(quantizer).rcp, stream=stream)
        if self._settled_host is not None and not self._settled_seen:
            if stream is None:
                self._settled_host.copy_(self._settled, non_blocking=True)
            else:
                with torch.cuda.stream(stream):
                    self._settled_host.copy_(self._settled, non_blocking=True)
        if stream is None:
            if self._eager:
                self.check_finite()
            self._nsteps += 1
        return codes, rowsum

    # ---- overlapped parameter steps -------------------------------------------------------------------------
    def _issue_ahead(self) -> None:
        """Launch this step's fused kernel on the block's side stream, ahead of the call that will use it.  Only for
        a Parameter that has not changed since this step's previous launch: the running range already covers those
        values, so the update is idempotent whatever happens to the result -- if the tensor is modified before the
        result is used, the result is dropped and the step runs again in line."""
        st, quantizer = self._state, self._quantizer_ref()
        if st is None or quantizer is None or self._pdata is None or torch.is_grad_enabled():
            return
        ref, version, mode = self._pdata
        data = ref()
        if data is None or data._version != version or self._alias_of is not None or self.min is None \
                or quantizer.has_uninitialized_params or self._eager or _host_of(quantizer).get_export_mode():
            return
        cur = torch.cuda.current_stream(data.device)
        side = st.side_stream(data.device)
        fork = torch.cuda.Event()
        fork.record(cur)
        side.wait_event(fork)
        codes, rowsum = self._launch(quantizer, data, mode, stream=side)
        done = torch.cuda.Event()
        done.record(side)
        self._ahead = (done, ref, version, (codes, rowsum))

    def _take_ahead(self, data: torch.Tensor):
        """(codes, rowsum) of the launch in flight for ``data`` -- the current stream now waits for it -- or None."""
        if self._ahead is None:
            return None
        done, ref, version, result = self._ahead
        self._ahead = None
        torch.cuda.current_stream(data.device if isinstance(data, torch.Tensor) and data.is_cuda else None).wait_event(done)
        if ref() is not data or data._version != version:
            return None
        return result

    def _drop_ahead(self) -> None:
        if self._ahead is not None:
            done = self._ahead[0]
            self._ahead = None
            torch.cuda.current_stream(self.min.device if self.min is not None else None).wait_event(done)

    def forward(self, quantizer, callback, args: tuple, kwargs: dict):
        data = args[0] if args else kwargs.get("data", next(iter(kwargs.values()), None))
        mode = self._fused_mode(quantizer, callback, data)
        if mode and self.min is not None and self._alias_of is None and \
                (self.min.device != data.device or torch.promote_types(self.min.dtype, data.dtype) != self.min.dtype):
            mode = 0     # a continued range on another device / in a narrower dtype: the plain path converts it
        if not mode:
            self._drop_ahead()
            return super().forward(quantizer, callback, args, kwargs)
        if not self._initialized:
            self.setup_estimator(data)
            self._initialized = True
        return self._fused_step(quantizer, data, mode)

    def extra_repr(self) -> str:
        return f"min={self.min}, max={self.max}"


def _unalias_prepare(steps):
    """First half of giving aliased quantizers their own parameter storage back when the block ends: the new tensors
    are allocated and installed (host work that needs nothing from the device), the values follow with
    ``_unalias_finish`` -- after whatever still writes the shared storage (the re-derivation from merged ranges).
    The running ranges stay shared views: the steps end with the block."""
    dsts, srcs = [], []
    for s in steps:
        if s._alias_of is None:
            continue
        s._alias_of = None
        q = s._quantizer_ref()
        if q is None:
            continue
        with torch.no_grad():
            for p in (q.scale, q.offset):
                if p is not None:
                    src = p.data
                    dst = torch.empty_like(src)
                    p.data = dst
                    dsts.append(dst)
                    srcs.append(src)
    return dsts, srcs


def _unalias_finish(plan) -> None:
    """Second half: ONE multi-tensor copy per (device, dtype) fills the storage ``_unalias_prepare`` installed."""
    dsts, srcs = plan
    if not dsts:
        return
    groups: dict = {}
    for d, s_ in zip(dsts, srcs):
        groups.setdefault((d.device, d.dtype), ([], []))
        groups[(d.device, d.dtype)][0].append(d)
        groups[(d.device, d.dtype)][1].append(s_)
    with torch.no_grad():
        for d, s_ in groups.values():
            torch._foreach_copy_(d, s_)


def _unalias_all(steps) -> None:
    """Both halves at once, for when nothing will rewrite the shared storage any more: the copies are allocated AND
    filled by one multi-tensor op per (device, dtype) (``x * 1`` is exact and keeps the sign of zero and NaN) instead
    of one Python-level allocation per tensor."""
    owners, srcs = [], []
    for s in steps:
        if s._alias_of is None:
            continue
        s._alias_of = None
        q = s._quantizer_ref()
        if q is None:
            continue
        for p in (q.scale, q.offset):
            if p is not None:
                owners.append(p)
                srcs.append(p.data)
    groups: dict = {}
    for owner, src in zip(owners, srcs):
        groups.setdefault((src.device, src.dtype), []).append((owner, src))
    with torch.no_grad():
        for ents in groups.values():
            tensors = [src for _, src in ents]
            copies = torch._foreach_mul(tensors, 1) if tensors[0].is_floating_point() and len(tensors) > 1 \
                else [t.clone() for t in tensors]
            for (owner, _), c in zip(ents, copies):
                owner.data = c


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
        seen = set()
        for _, quantizer in module.named_modules(remove_duplicate=True):
            if not _is_quantizer(quantizer) or quantizer.is_stub() or id(quantizer) in seen:
                continue
            seen.add(id(quantizer))
            if isinstance(quantizer, RangeSettable) or not self.skip_unsupported_quantizers:
                yield quantizer
            else:
                logger.warning(f"{type(quantizer).__name__} does not implement RangeSettable. Therefore it is not "
                               f"included in {type(self).__name__} range setting.")


class RunningMinMaxRangeEstimator(_MinMaxRangeEstimatorBase):
    """``ff.estimate_ranges(model, running_minmax)``.  Keyword extras (all default to the reference's behaviour where
    they change anything observable):

    eager_checks        raise ``NotImplementedError("Infinite")`` at the offending step (one host sync per step)
                        instead of when the block ends
    sync_ranges /       data-parallel calibration: all-reduce (MIN/MAX) the running ranges of the quantizers whose
    process_group       input is NOT an ``nn.Parameter`` when the block ends (parameters are rank-identical replicas;
                        ``sync_parameters=True`` reduces them too)
    fused               one kernel per quantizer per forward where the layout allows (default True)
    dedupe              identically configured quantizers fed the same tensor object share one launch (default True)
    memoize_parameters  an unchanged ``nn.Parameter`` is not re-quantized on later steps (default True)
    overlap_parameters  N > 0 (and ``memoize_parameters=False``): from the second step on, the fused step of a weight is
                        launched N weight quantizers ahead of its use on a side stream, so the HBM-bound kernel runs
                        under the tensor-bound linears of the layers before it (a parallel branch when the step is
                        captured in a CUDA graph).  Same work per step, same results: only an unchanged Parameter
                        is launched ahead, for which the running-range update is idempotent
    """

    def __init__(self, disable_quantization: bool = False, skip_unsupported_quantizers: bool = False, *,
                 eager_checks: bool = False, sync_ranges: bool = False, process_group=None, fused: bool = True,
                 dedupe: bool = True, memoize_parameters: bool = True, sync_parameters: bool = False,
                 overlap_parameters: int = 0) -> None:
        self.disable_quantization = disable_quantization
        self.skip_unsupported_quantizers = skip_unsupported_quantizers
        self.eager_checks = eager_checks
        self.sync_ranges = sync_ranges or process_group is not None
        self.process_group = process_group
        self.fused = fused
        self.dedupe, self.memoize_parameters, self.sync_parameters = dedupe, memoize_parameters, sync_parameters
        self.overlap_parameters = int(overlap_parameters)
        self._steps: list = []
        self._state = _BlockState(dedupe, memoize_parameters, overlap_parameters)
        self.last_stats: dict = {}
        self.last_exit: dict = {}

    def prepare(self, module):
        self._check(module)
        step = RunningMinMaxEstimator(module, disable_quantization=self.disable_quantization,
                                      eager_checks=self.eager_checks, fused=self.fused, state=self._state)
        self._steps.append((module, step))
        return module.register_override(step)

    def cleanup(self, module, metadata) -> None:
        # under the reference's own estimate_ranges nobody calls finalize(): the first cleanup of a block that ended
        # without an exception does
        if self._steps:
            if sys.exc_info()[0] is None:
                self.finalize(())
            else:                                # the block failed: drop its half-finished state
                self._state.drain()
                for s in self._state.aliases:
                    s._unalias()
                self._steps, self._state = [], _BlockState(self.dedupe, self.memoize_parameters, self.overlap_parameters)
        super().cleanup(module, metadata)

    # ---- block exit -------------------------------------------------------------------------------------
    finalize_runs_cleanup = True      # estimate_ranges hands its cleanup loop to finalize(), see below

    def finalize(self, prepared: Sequence[tuple], cleanup=None) -> None:
        """The block ends.  ONE host sync, and everything the host can do without the device's answer happens BEFORE
        it, while the device is still working through the queued steps: the slot signature and its all-reduce, the
        descriptor table of the batched range -> parameters launch, fresh storage for the de-duplicated quantizers,
        and (``cleanup``, passed by this package's ``estimate_ranges``) the removal of the overrides.  After the sync:
        two all-reduces, one parameters launch, one multi-tensor copy."""
        del prepared
        import time

        steps, state = self._steps, self._state
        self._steps, self._state = [], _BlockState(self.dedupe, self.memoize_parameters, self.overlap_parameters)
        t0 = time.perf_counter()
        state.drain()
        exchange = None
        unalias = ([], [])
        flags_value = None
        detail: dict = {}
        try:
            if self.sync_ranges:
                exchange = self._sync_begin(steps, state)
            if exchange is None:
                _unalias_all(state.aliases)     # nothing will rewrite the shared storage: copy now, before the sync
            else:
                unalias = _unalias_prepare(state.aliases)
            state.recent.clear()
            # the reference's per-step `isinf().any()` checks, folded into one read; a data-parallel exit carries the
            # flags of the block's arenas inside the signature message instead
            flags = [s.flags for _, s in steps if s.flags is not None and s._state is None]
            if exchange is None:
                flags = state.all_flags() + flags
            pending = []                        # (host-visible tensor, device to synchronise or None), one per device
            if flags and not self.eager_checks:
                by_device: dict = {}
                for f in {id(f): f for f in flags}.values():
                    by_device.setdefault(f.device, []).append(f.reshape(()))
                for device, fs in by_device.items():
                    stacked = torch.stack(fs)
                    if stacked.is_cuda:
                        host = torch.empty(stacked.shape, dtype=stacked.dtype).pin_memory()
                        host.copy_(stacked, non_blocking=True)
                        pending.append((host, device))
                    else:
                        pending.append((stacked, None))
            if cleanup is not None:
                cleanup()
            t1 = time.perf_counter()
            # ---- the exit's one host sync -----------------------------------------------------------------
            if exchange is not None:
                flags_value = self._sync_finish(exchange, state, detail)
            t2 = time.perf_counter()
        finally:
            _unalias_finish(unalias)
            _unalias_all(state.aliases)         # only after a failure above: whoever is still aliased gets its own storage
        value = flags_value or 0
        for host, device in pending:
            if device is not None:
                # single process: THE host sync of the exit; data parallel: the signature read has drained this stream already
                torch.cuda.current_stream(device).synchronize()
            for v in host.tolist():
                value |= int(v)
        t3 = time.perf_counter()
        self.last_stats = dict(state.stats)
        self.last_exit = {"before_sync_ms": (t1 - t0) * 1e3, "sync_and_exchange_ms": (t2 - t1) * 1e3,
                          "flags_ms": (t3 - t2) * 1e3, "aliased": len(state.aliases), **detail}
        if not self.eager_checks or flags_value is not None:
            _raise_for_flags(value)

    def _sync_begin(self, steps, state):
        """Data-parallel exit, the part before the host sync.  Slots of the activation arena are handed out in the
        order quantizers first see data, which control flow that depends on the data (experts without tokens on one
        shard) can make rank dependent: the ranks first agree on the slot layout with one small MAX all-reduce of a
        signature vector, which also carries the block's +-inf flags.  Issued here, read in ``_sync_finish``; in
        between the host prepares the batched parameters launch for the common case of identical layouts."""
        import torch.distributed as dist

        from ..distributed import _active

        group = self.process_group
        if not _active(group):
            return None
        todo = [(i, q, s) for i, (q, s) in enumerate(steps)
                if s._alias_of is None and (self.sync_parameters or not s._param_data)]
        device = None
        for _, _, s in todo:
            if s.min is not None:
                device = s.min.device
                break
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() and \
                dist.get_backend(group) == "nccl" else torch.device("cpu")
        # signature: slot position inside the activation arena, -1 = not seen here, -2 = seen but outside the arena
        sig = []
        for _, _, s in todo:
            if s.min is None:
                sig.append(-1)
            elif s._slot is not None and s.min.numel() == s._slot[3] and (s._slot[0] is state.act or s._slot[0] is state.param):
                sig.append((int(s._slot[0] is state.param) << 60) | (s._slot[1] << 40) | s._slot[2])
            else:
                sig.append(-2)
        n = len(sig)
        # [sig | -sig | flag bit 0 | flag bit 1]: MAX over the ranks gives max(sig), -min(sig) and the OR of each flag bit
        msg = torch.tensor(sig + [-v for v in sig] + [0, 0], dtype=torch.int64)
        if device.type == "cuda":
            msg = msg.pin_memory()
        msg = msg.to(device, non_blocking=True)
        flags = [f for f in state.all_flags() if f.device == device]
        if flags:
            f64 = torch.stack([f.reshape(()) for f in flags]).to(torch.int64)
            msg[2 * n:] |= torch.stack([(f64 & 1).amax(), ((f64 >> 1) & 1).amax()])
        dist.all_reduce(msg, op=dist.ReduceOp.MAX, group=group)
        if msg.is_cuda:
            got = torch.empty(msg.shape, dtype=msg.dtype).pin_memory()
            got.copy_(msg, non_blocking=True)
        else:
            got = msg
        live = [(q, s) for _, q, s in todo if s.min is not None]
        launches = self._prepare_parameters_from_ranges(live, state) if -2 not in sig else None
        return {"todo": todo, "sig": sig, "device": device, "group": group, "got": got, "live": live, "launches": launches}

    def _sync_finish(self, ex, state, detail: dict) -> int:
        import time

        from ..distributed import all_reduce_minmax_buffers

        t0 = time.perf_counter()
        device, group, sig, n = ex["device"], ex["group"], ex["sig"], len(ex["sig"])
        if device.type == "cuda":
            torch.cuda.current_stream(device).synchronize()     # the exit's one host sync
        got = ex["got"].tolist()
        t1 = time.perf_counter()
        same = all(a == -b for a, b in zip(got[:n], got[n:2 * n])) and -2 not in sig
        flags_value = int(got[2 * n]) | (int(got[2 * n + 1]) << 1)
        other_flags = [f for f in state.all_flags() if f.device != device]     # a block spanning devices: the old way
        if same:
            all_reduce_minmax_buffers(list(state.act.buffers()), other_flags, group=group)
            if self.sync_parameters:
                all_reduce_minmax_buffers(list(state.param.buffers()), None, group=group)
        else:
            self._sync_packed(ex["todo"], device, group)
            all_reduce_minmax_buffers([], other_flags, group=group)
        for f in other_flags:
            flags_value |= int(f.item())
        t2 = time.perf_counter()
        if same and ex["launches"] is not None:
            for launch in ex["launches"]:
                launch()
        else:
            for launch in self._prepare_parameters_from_ranges([(q, s) for q, s in ex["live"] if s.min is not None], state):
                launch()
        t3 = time.perf_counter()
        detail.update({"wait_ms": (t1 - t0) * 1e3, "allreduce_issue_ms": (t2 - t1) * 1e3, "params_ms": (t3 - t2) * 1e3,
                       "exchanged_quantizers": n, "same_layout": same})
        return flags_value

    def _sync(self, steps, state) -> Optional[int]:
        """Both halves back to back (tests; the block exit interleaves its other host work between them)."""
        ex = self._sync_begin(steps, state)
        return None if ex is None else self._sync_finish(ex, state, {})

    def _sync_packed(self, todo, device, group) -> None:
        import torch.distributed as dist

        sizes = torch.tensor([0 if s.min is None else s.min.numel() for _, _, s in todo], dtype=torch.int64, device=device)
        lo_sz = torch.where(sizes > 0, sizes, torch.full_like(sizes, torch.iinfo(torch.int64).max))
        hi_sz = sizes.clone()
        dist.all_reduce(hi_sz, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(lo_sz, op=dist.ReduceOp.MIN, group=group)
        hi, lo = hi_sz.tolist(), lo_sz.tolist()
        for (i, q, s), a, b in zip(todo, hi, lo):
            if a > 0 and b != a:
                raise RuntimeError(f"estimate_ranges: quantizer #{i} ({type(q).__name__}) has {b} range entries on one rank "
                                   f"and {a} on another; data-parallel calibration needs identical granularities")
        total = sum(hi)
        if total == 0:
            return
        packed_min = torch.full((total,), float("inf"), dtype=torch.float32, device=device)
        packed_max = torch.full((total,), float("-inf"), dtype=torch.float32, device=device)
        pos = 0
        for (_, _, s), n in zip(todo, hi):
            if s.min is not None:
                packed_min[pos:pos + n] = s.min.reshape(-1).float()
                packed_max[pos:pos + n] = s.max.reshape(-1).float()
            pos += n
        dist.all_reduce(packed_min, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(packed_max, op=dist.ReduceOp.MAX, group=group)
        pos = 0
        for (_, q, s), n in zip(todo, hi):
            if n and s.min is not None:
                s.min.copy_(packed_min[pos:pos + n].to(s.min.dtype))
                s.max.copy_(packed_max[pos:pos + n].to(s.max.dtype))
            elif n:
                # this rank never ran the quantizer: it adopts the range the other ranks measured
                s.min, s.max = packed_min[pos:pos + n].clone(), packed_max[pos:pos + n].clone()
                if torch.isfinite(s.min).all() and not _set_range_direct(q, s.min, s.max):
                    q.quantization_range = (s.min, s.max)
                s.min = None            # already applied
            pos += n

    def _prepare_parameters_from_ranges(self, steps, state) -> list:
        """After the ranges were merged across ranks every quantizer's (scale, offset) follows from its merged range.
        Returns the launches as closures, so that the tables can be built before the exit's host sync and fired
        after the all-reduces were issued: quantizers whose running range lives in an arena chunk and whose
        parameters are materialised fp32 tensors of the right size share ONE launch per chunk (descriptor table
        already on the device); the rest go through the ``quantization_range`` setter."""
        per_chunk: dict = {}
        launches: list = []

        def through_setter(quantizer, step):
            def run():
                if step.min is not None and not _set_range_direct(quantizer, step.min, step.max):
                    quantizer.quantization_range = (step.min, step.max)
            return run

        for quantizer, step in steps:
            slot = step._slot
            entry = None
            if slot is not None and not quantizer.has_uninitialized_params:
                scale, offset = quantizer.scale, quantizer.offset
                n = slot[3]
                if scale.dtype == torch.float32 and scale.numel() == n and scale.is_contiguous() and \
                        (offset is None or (offset.dtype == torch.float32 and offset.numel() == n and offset.is_contiguous())):
                    entry = (slot[2], n, quantizer.num_bits, quantizer.symmetric, quantizer.allow_one_sided,
                             scale.data, None if offset is None else offset.data)
            if entry is None:
                launches.append(through_setter(quantizer, step))
                continue
            per_chunk.setdefault((id(slot[0]), slot[1], step.min.device, step.min.dtype, _host_of(quantizer).rcp),
                                 (slot, []))[1].append(entry)
        for (_, ci, dev, dt, rcp), (slot, entries) in per_chunk.items():
            mn, mx, used = slot[0].chunks[(dev, dt)][ci]
            launches.append(ops.parameters_for_ranges_batched_prepare(mn, mx, entries, reciprocal_scalar_division=rcp))
        return launches


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
