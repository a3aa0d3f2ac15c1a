"""Tensor-level entry points of the hot path: the four ops the reference defines as
``torch.ops.fastforward.*`` (quantization/_quantizer_impl.py:144,172,193,243), with the same
argument order, defaults, return conventions and error types -- plus the fused / sync-free
extras the B200 backend adds (fake_quantize_by_tile, tile_minmax, running_minmax_update_,
parameters_for_range_).

Everything here is glue: argument validation, output allocation on the input's device and
current stream, and one C-ABI call.  No arithmetic happens in Python.
"""

from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi as C
from .exceptions import QuantizationError

_MANTISSA = {torch.bfloat16: 7, torch.float16: 10, torch.float32: 23, torch.float64: 52}


def can_support_bitwidth(dtype: torch.dtype, num_bits: float) -> bool:
    """quantization/_quantizer_impl.py:44-75."""
    if dtype.is_floating_point or dtype.is_complex:
        if dtype in _MANTISSA:
            avail = _MANTISSA[dtype]
        elif dtype in (torch.float8_e4m3fn, torch.float8_e4m3fnuz):
            avail = 3
        elif dtype in (torch.float8_e5m2, torch.float8_e5m2fnuz):
            avail = 2
        else:
            avail = num_bits
    else:
        avail = torch.iinfo(dtype).bits
    return avail + 2 >= num_bits


def _on_device(fn):
    """Run the op with its first tensor argument's device current (see _cabi.device_of): tensors on a GPU other than
    the current one work, as they do in the reference (device_map-sharded models)."""
    import functools

    @functools.wraps(fn)
    def wrapped(first, *args, **kwargs):
        guard = C.device_of(first.device) if isinstance(first, torch.Tensor) and first.is_cuda else C._NO_GUARD
        if guard is C._NO_GUARD:
            return fn(first, *args, **kwargs)
        with guard:
            return fn(first, *args, **kwargs)
    return wrapped


def _flags(allow_one_sided, reciprocal_scalar_division: bool = False) -> int:
    """The C ABI's `allow_one_sided` word: bit 0 the flag itself, bit 1 FFQ_FLAG_SCALAR_DIV_RECIPROCAL (the scalar
    divisions of parameters_for_range done as aten's CUDA kernel does them: what the reference computes on a GPU)."""
    return int(bool(allow_one_sided)) | (2 if reciprocal_scalar_division else 0)


def _bitwidth_guard(dtype: torch.dtype, num_bits: float) -> None:
    if not can_support_bitwidth(dtype, num_bits):
        raise RuntimeError(f"Provided dtype ({dtype}) is not enough to store {num_bits} bits quantized values.")


def _tile(data: torch.Tensor, tile_size) -> tuple:
    if isinstance(tile_size, str):  # "data_shape"
        return tuple(data.shape)
    return tuple(int(t) for t in tile_size)


def _num_tiles(shape: Sequence[int], tile: Sequence[int]) -> int:
    n = 1
    for d, t in zip(shape, tile):
        n *= d // t if t else 0
    return n


def _param(p: Optional[torch.Tensor], ntiles: int, device: torch.device, what: str) -> Optional[torch.Tensor]:
    """Flatten a parameter; a one-element parameter broadcasts over the tiles like the
    reference's ``scale[:, None]`` does; any other size mismatch is the reference's RuntimeError."""
    if p is None:
        return None
    if p.dim() == 1 and p.numel() == ntiles and p.device == device and p.is_contiguous():
        return p          # the common case: nothing to reshape, expand or copy
    if p.device != device:
        raise RuntimeError(f"Expected all tensors to be on the same device, but '{what}' is on {p.device} and data on {device}")
    p = p.detach().reshape(-1)
    if p.numel() != ntiles:
        if p.numel() == 1:
            p = p.expand(ntiles)
        else:
            raise RuntimeError(
                f"The size of '{what}' ({p.numel()}) must match the number of tiles ({ntiles})"
            )
    return p.contiguous()


def _prep(data: torch.Tensor, tile_size, what: str = "data"):
    C.require_cuda(data, what)
    tile = _tile(data, tile_size)
    shape = tuple(data.shape)
    # an empty tensor is returned untouched by the reference before any tile check (tiled_tensor.py:85-86)
    layout = C.make_layout(shape, tile) if data.numel() else None
    return (data if data.is_contiguous() else data.contiguous()), shape, tile, layout


# ------------------------------------------------------------------------------------------
@_on_device
def quantize_by_tile(
    data: torch.Tensor,
    scale: torch.Tensor,
    tile_size,
    num_bits: float,
    output_dtype: Optional[torch.dtype],
    offset: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """``torch.ops.fastforward.quantize_by_tile`` (quantization/_quantizer_impl.py:144-169)."""
    x, shape, tile, layout = _prep(data, tile_size)
    promoted = torch.promote_types(x.dtype, scale.dtype)
    if offset is not None:
        promoted = torch.promote_types(promoted, offset.dtype)
    out_dtype = output_dtype or promoted
    _bitwidth_guard(out_dtype, num_bits)
    q = torch.empty(shape, dtype=out_dtype, device=x.device)
    if x.numel() == 0:
        return q
    nt = layout.num_tiles
    s = _param(scale, nt, x.device, "scale")
    o = _param(offset, nt, x.device, "offset")
    C.check(C.lib.ffq_quantize(
        x.data_ptr(), C.dtype_tag(x.dtype), q.data_ptr(), C.dtype_tag(out_dtype),
        s.data_ptr(), C.dtype_tag(s.dtype), C.ptr(o), C.dtype_tag(o.dtype if o is not None else None),
        layout.ref, float(num_bits), C.current_stream(x.device)))
    return q


@_on_device
def dequantize_by_tile(
    data: torch.Tensor,
    scale: torch.Tensor,
    tile_size,
    offset: Optional[torch.Tensor] = None,
    output_dtype: Optional[torch.dtype] = None,
) -> torch.Tensor:
    """``torch.ops.fastforward.dequantize_by_tile`` (quantization/_quantizer_impl.py:172-190)."""
    q, shape, tile, layout = _prep(data, tile_size)
    promoted = torch.promote_types(q.dtype, offset.dtype if offset is not None else scale.dtype)
    promoted = torch.promote_types(promoted, scale.dtype)
    out_dtype = output_dtype or promoted
    y = torch.empty(shape, dtype=out_dtype, device=q.device)
    if q.numel() == 0:
        return y
    nt = layout.num_tiles
    s = _param(scale, nt, q.device, "scale")
    o = _param(offset, nt, q.device, "offset")
    C.check(C.lib.ffq_dequantize(
        q.data_ptr(), C.dtype_tag(q.dtype), y.data_ptr(), C.dtype_tag(out_dtype),
        s.data_ptr(), C.dtype_tag(s.dtype), C.ptr(o), C.dtype_tag(o.dtype if o is not None else None),
        layout.ref, C.current_stream(q.device)))
    return y


@_on_device
def fake_quantize_by_tile(
    data: torch.Tensor,
    scale: torch.Tensor,
    tile_size,
    num_bits: float,
    quantized_dtype: Optional[torch.dtype] = None,
    offset: Optional[torch.Tensor] = None,
    output_dtype: Optional[torch.dtype] = None,
    return_codes: bool = False,
):
    """Fused ``dequantize_by_tile(quantize_by_tile(x))`` in one pass over HBM -- bit-identical to
    the two-op sequence of affine/function.py:94-121 / quantization/fuse.py:91-121."""
    x, shape, tile, layout = _prep(data, tile_size)
    q_dtype = quantized_dtype or x.dtype
    _bitwidth_guard(q_dtype, num_bits)
    out_dtype = output_dtype or (x.dtype if x.dtype.is_floating_point else scale.dtype)
    y = torch.empty(shape, dtype=out_dtype, device=x.device)
    codes = torch.empty(shape, dtype=q_dtype, device=x.device) if return_codes else None
    if x.numel() > 0:
        nt = layout.num_tiles
        s = _param(scale, nt, x.device, "scale")
        o = _param(offset, nt, x.device, "offset")
        C.check(C.lib.ffq_fakequant_fwd(
            x.data_ptr(), C.dtype_tag(x.dtype), y.data_ptr(), C.dtype_tag(out_dtype),
            C.ptr(codes), C.dtype_tag(q_dtype),
            s.data_ptr(), C.dtype_tag(s.dtype), C.ptr(o), C.dtype_tag(o.dtype if o is not None else None),
            layout.ref, float(num_bits), C.current_stream(x.device)))
    return (y, codes) if return_codes else y


@_on_device
def quantize_by_tile_backward(
    data: torch.Tensor,
    output_grad: torch.Tensor,
    scale: torch.Tensor,
    tile_size,
    num_bits: float,
    offset: Optional[torch.Tensor] = None,
) -> List[torch.Tensor]:
    """``torch.ops.fastforward.quantize_by_tile_backward`` (quantization/_quantizer_impl.py:193-237):
    returns ``[dx, dscale, doffset]``; ``doffset`` is an empty tensor when ``offset`` is None."""
    x, shape, tile, layout = _prep(data, tile_size)
    C.require_cuda(output_grad, "output_grad")
    g = output_grad.detach().contiguous()
    if tuple(g.shape) != shape:
        raise RuntimeError(f"output_grad shape {tuple(g.shape)} does not match data shape {shape}")
    dx = torch.empty(shape, dtype=g.dtype, device=x.device)
    dscale = torch.empty(scale.shape, dtype=scale.dtype, device=x.device)
    doff_dtype = torch.promote_types(scale.dtype, g.dtype)
    doffset = torch.empty(scale.shape, dtype=doff_dtype, device=x.device) if offset is not None else torch.Tensor()
    if x.numel() == 0:
        dscale.zero_()
        if offset is not None:
            doffset.zero_()
        return [dx, dscale, doffset]
    nt = layout.num_tiles
    s = _param(scale, nt, x.device, "scale")
    o = _param(offset, nt, x.device, "offset")
    if dscale.numel() != nt:  # one-element scale broadcast over many tiles: reduce afterwards
        dscale_full = torch.empty(nt, dtype=scale.dtype, device=x.device)
        doffset_full = torch.empty(nt, dtype=doff_dtype, device=x.device) if offset is not None else None
    else:
        dscale_full, doffset_full = dscale, (doffset if offset is not None else None)
    ws_bytes = C.workspace_bytes(C.WS_QUANTIZE_BWD, layout, x.dtype)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
    C.check(C.lib.ffq_quantize_bwd(
        x.data_ptr(), C.dtype_tag(x.dtype), g.data_ptr(), C.dtype_tag(g.dtype), dx.data_ptr(),
        dscale_full.data_ptr(), C.ptr(doffset_full),
        s.data_ptr(), C.dtype_tag(s.dtype), C.ptr(o), C.dtype_tag(o.dtype if o is not None else None),
        layout.ref, float(num_bits), C.ptr(ws), ws_bytes, C.current_stream(x.device)))
    if dscale_full is not dscale:
        dscale.copy_(dscale_full.sum().reshape(scale.shape))
        if offset is not None:
            doffset.copy_(doffset_full.sum().reshape(scale.shape))
    return [dx, dscale, doffset]


@_on_device
def quantize_dynamic_by_tile(
    data: torch.Tensor,
    tile_size,
    num_bits: float,
    symmetric: bool,
    allow_one_sided: bool,
    output_dtype: Optional[torch.dtype],
    reciprocal_scalar_division: bool = False,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``torch.ops.fastforward.quantize_dynamic_by_tile`` (quantization/_quantizer_impl.py:243-285):
    returns ``(codes, scale, rounded_offset)``, scale/offset in float32."""
    x, shape, tile, layout = _prep(data, tile_size)
    if x.numel() == 0:
        raise QuantizationError(f"Cannot dynamically quantize an empty tensor of shape {data.shape}")
    out_dtype = output_dtype or torch.promote_types(x.dtype, torch.float32)
    _bitwidth_guard(out_dtype, num_bits)
    nt = _num_tiles(shape, tile)
    q = torch.empty(shape, dtype=out_dtype, device=x.device)
    scale = torch.empty(nt, dtype=torch.float32, device=x.device)
    offset = torch.empty(nt, dtype=torch.float32, device=x.device)
    ws_bytes = C.workspace_bytes(C.WS_DYNAMIC_QUANTIZE, layout, x.dtype)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    C.check(C.lib.ffq_dynamic_quantize(
        x.data_ptr(), C.dtype_tag(x.dtype), q.data_ptr(), C.dtype_tag(out_dtype),
        scale.data_ptr(), offset.data_ptr(), layout.ref, float(num_bits),
        int(bool(symmetric)), _flags(allow_one_sided, reciprocal_scalar_division), ws.data_ptr(), ws_bytes,
        C.current_stream(x.device)))
    return q, scale, offset


# ------------------------------------------------------------------------------------------
# range estimation pieces (range_setting/minmax.py:226-237, quantization/affine/range.py:54-122)
# ------------------------------------------------------------------------------------------
def _minmax_call(x, layout, tile_min, tile_max, run_min, run_max, flags):
    ws_bytes = C.workspace_bytes(C.WS_MINMAX, layout, x.dtype)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
    C.check(C.lib.ffq_minmax(
        x.data_ptr(), C.dtype_tag(x.dtype), C.ptr(tile_min), C.ptr(tile_max), C.ptr(run_min), C.ptr(run_max),
        C.dtype_tag(run_min.dtype if run_min is not None else None),
        C.ptr(flags), layout.ref, C.ptr(ws), ws_bytes, C.current_stream(x.device)))


@_on_device
def tile_minmax(data: torch.Tensor, tile_size) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-tile ``(min, max)`` in the data dtype: ``torch.min/max(tiles_to_rows(data), -1).values``."""
    x, shape, tile, layout = _prep(data, tile_size)
    if x.numel() == 0:
        raise IndexError("min(): Expected reduction dim 1 to have non-zero size.")
    nt = _num_tiles(shape, tile)
    mn = torch.empty(nt, dtype=x.dtype, device=x.device)
    mx = torch.empty(nt, dtype=x.dtype, device=x.device)
    _minmax_call(x, layout, mn, mx, None, None, None)
    return mn, mx


@_on_device
def running_minmax_update_(
    run_min: torch.Tensor, run_max: torch.Tensor, data: torch.Tensor, tile_size,
    flags: Optional[torch.Tensor] = None,
) -> None:
    """In-place ``run_min = min(run_min, tile_min(data))`` / ``run_max = max(...)`` in ONE pass
    over ``data``; ``flags`` (int32[1]) gets bit 0 set if a tile extremum is +-inf
    (range_setting/minmax.py:229-237, without its host sync)."""
    x, shape, tile, layout = _prep(data, tile_size)
    nt = _num_tiles(shape, tile)
    for t, name in ((run_min, "run_min"), (run_max, "run_max")):
        C.require_cuda(t, name)
        if t.numel() != nt or not t.is_contiguous() or t.dtype != run_min.dtype or \
                torch.promote_types(t.dtype, x.dtype) != t.dtype:
            raise RuntimeError(f"{name} must be a contiguous tensor with {nt} elements whose dtype holds {x.dtype}")
    if flags is not None and (flags.dtype != torch.int32 or not flags.is_cuda):
        raise RuntimeError("flags must be a CUDA int32 tensor")
    _minmax_call(x, layout, None, None, run_min, run_max, flags)


@_on_device
def parameters_for_range_(
    min_range: torch.Tensor, max_range: torch.Tensor, num_bits: float, symmetric: bool, allow_one_sided: bool,
    scale_out: torch.Tensor, offset_out: Optional[torch.Tensor], round_offset: bool = False,
    reciprocal_scalar_division: bool = False,
) -> None:
    """Device-side, sync-free ``parameters_for_range`` (quantization/affine/range.py:54-122) writing
    straight into a quantizer's ``scale`` / ``offset`` storage (nn/linear_quantizer.py:347-357)."""
    C.require_cuda(min_range, "min_range")
    mn = min_range if (min_range.dim() == 1 and min_range.is_contiguous()) else min_range.detach().reshape(-1).contiguous()
    mx = max_range if (max_range.dim() == 1 and max_range.is_contiguous()) else max_range.detach().reshape(-1).contiguous()
    if mn.dtype != mx.dtype:
        common = torch.promote_types(mn.dtype, mx.dtype)
        mn, mx = mn.to(common), mx.to(common)
    n = mn.numel()
    if mx.numel() != n or scale_out.numel() != n or (offset_out is not None and offset_out.numel() != n):
        raise RuntimeError("parameters_for_range_: min, max, scale and offset must have the same number of elements")
    if not scale_out.is_contiguous() or (offset_out is not None and not offset_out.is_contiguous()):
        raise RuntimeError("parameters_for_range_: outputs must be contiguous")
    ws = C.scratch(mn.device, 4096)
    C.check(C.lib.ffq_params_for_range(
        mn.data_ptr(), mx.data_ptr(), C.dtype_tag(mn.dtype), n, float(num_bits),
        int(bool(symmetric)), _flags(allow_one_sided, reciprocal_scalar_division), int(bool(round_offset)),
        scale_out.data_ptr(), C.dtype_tag(scale_out.dtype),
        C.ptr(offset_out), C.dtype_tag(offset_out.dtype if offset_out is not None else None),
        ws.data_ptr(), ws.numel(), C.current_stream(mn.device)))


@_on_device
def calibrate_fake_quantize_(
    data: torch.Tensor, tile_size, num_bits: float, symmetric: bool, allow_one_sided: bool,
    scale_out: torch.Tensor, offset_out: Optional[torch.Tensor], quantized_dtype: Optional[torch.dtype] = None,
    out: Optional[torch.Tensor] = None, run_min: Optional[torch.Tensor] = None, run_max: Optional[torch.Tensor] = None,
    flags: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
    reciprocal_scalar_division: bool = False,
) -> torch.Tensor:
    """Calibrate on ``data`` and snap it to the grid in ONE pass: per-tile min/max (merged into ``run_min`` /
    ``run_max`` when given) -> ``scale_out`` / ``offset_out`` in place -> ``out = dequantize(quantize(data))``
    (``out=data`` for in place).  Equals ``tile_minmax`` + ``parameters_for_range_`` + ``fake_quantize_by_tile`` bit for
    bit.  Raises NotImplementedError for layouts the fused kernels do not cover (calibrate_quantize_mode not 1 or 3)."""
    x, shape, tile, layout = _prep(data, tile_size)
    code_dtype = quantized_dtype or x.dtype
    _bitwidth_guard(code_dtype, num_bits)
    if x.numel() == 0:
        raise NotImplementedError("calibrate_fake_quantize_: empty tensor")
    nt = layout.num_tiles
    for t, name in ((scale_out, "scale_out"), (offset_out, "offset_out")):
        if t is not None and (t.dtype != torch.float32 or t.numel() != nt or not t.is_contiguous() or not t.is_cuda):
            raise RuntimeError(f"{name} must be a contiguous CUDA float32 tensor with {nt} elements")
    if (run_min is None) != (run_max is None):
        raise RuntimeError("run_min and run_max go together")
    for t, name in ((run_min, "run_min"), (run_max, "run_max")):
        if t is not None and (t.numel() != nt or not t.is_contiguous() or t.dtype != run_min.dtype or
                              torch.promote_types(t.dtype, x.dtype) != t.dtype):
            raise RuntimeError(f"{name} must be a contiguous tensor with {nt} elements whose dtype holds {x.dtype}")
    if out is None:
        out = torch.empty(shape, dtype=x.dtype, device=x.device)
    elif out.shape != x.shape or out.dtype != x.dtype or not out.is_contiguous() or out.device != x.device:
        raise RuntimeError("out must be a contiguous tensor like data")
    ws = workspace if workspace is not None else C.barrier_workspace(x.device, _CALQ_WS)
    C.check(C.lib.ffq_calibrate_fakequant(
        x.data_ptr(), C.dtype_tag(x.dtype), out.data_ptr(), C.ptr(run_min), C.ptr(run_max),
        C.dtype_tag(run_min.dtype if run_min is not None else None), scale_out.data_ptr(), C.ptr(offset_out), C.ptr(flags),
        layout.ref, float(num_bits), int(bool(symmetric)), _flags(allow_one_sided, reciprocal_scalar_division), C.dtype_tag(code_dtype),
        ws.data_ptr(), ws.numel(), C.current_stream(x.device)))
    return out


class FakeQuantBatch:
    """Device-side descriptor table of ``calibrate_fake_quantize_batched_``: built once for a set of (tensor, scale,
    offset) triples and reusable for as long as their storages stay where they are (the call itself is then free of
    host-to-device copies, so it can be captured in a CUDA graph)."""

    def __init__(self, tensors, scales, offsets, tile_len: int) -> None:
        dev = tensors[0].device
        rows, starts, blocks = [], [0], 0
        for w, s, o in zip(tensors, scales, offsets):
            nt = w.numel() // tile_len
            if w.device != dev or w.dtype != tensors[0].dtype or not w.is_contiguous() or w.numel() % tile_len or \
                    w.data_ptr() % 32 or s.dtype != torch.float32 or s.numel() != nt or not s.is_contiguous() or \
                    (o is not None and (o.dtype != torch.float32 or o.numel() != nt or not o.is_contiguous())):
                raise NotImplementedError("calibrate_fake_quantize_batched_: tensor outside the batched kernel's layout")
            rows.append((w.data_ptr(), w.data_ptr(), s.data_ptr(), 0 if o is None else o.data_ptr(), w.numel()))
            blocks += (nt + 255) // 256
            starts.append(blocks)
        self.key = tuple(r[:4] for r in rows)
        self.items = torch.tensor(rows, dtype=torch.int64).to(dev)
        self.block_start = torch.tensor(starts, dtype=torch.int64).to(torch.int32).to(dev)      # read as uint32
        self.workspace = torch.zeros(8 * len(rows), dtype=torch.uint8, device=dev)
        self.n, self.blocks, self.dtype, self.device, self.tile_len = len(rows), blocks, tensors[0].dtype, dev, tile_len


def calibrate_fake_quantize_batched_(batch: FakeQuantBatch, num_bits: float, symmetric: bool, allow_one_sided: bool,
                                     quantized_dtype: Optional[torch.dtype] = None,
                                     reciprocal_scalar_division: bool = False) -> None:
    """``calibrate_fake_quantize_(w, out=w)`` for every tensor of ``batch`` in ONE launch (+ one fix-up launch):
    per-tile min/max -> scale/offset in place -> the tensor snapped to its grid in place.  16-bit tensors, tiles of 64
    or 128 contiguous elements, one quantizer configuration for all of them."""
    code_dtype = quantized_dtype or batch.dtype
    _bitwidth_guard(code_dtype, num_bits)
    with C.device_of(batch.device):
        C.check(C.lib.ffq_calibrate_fakequant_batched(
            batch.items.data_ptr(), batch.block_start.data_ptr(), batch.n, batch.blocks, C.dtype_tag(batch.dtype),
            batch.tile_len, float(num_bits), int(bool(symmetric)), _flags(allow_one_sided, reciprocal_scalar_division),
            C.dtype_tag(code_dtype), batch.workspace.data_ptr(), batch.workspace.numel(), C.current_stream(batch.device)))


@_on_device
def parameters_for_ranges_batched_(min_buf: torch.Tensor, max_buf: torch.Tensor, entries,
                                   reciprocal_scalar_division: bool = False) -> None:
    """``parameters_for_range_`` for many quantizers in ONE launch.  ``min_buf`` / ``max_buf``: contiguous buffers
    holding every quantizer's running range; ``entries``: ``(start, length, num_bits, symmetric, allow_one_sided,
    scale, offset_or_None)`` with fp32 contiguous CUDA ``scale`` / ``offset`` of ``length`` elements."""
    parameters_for_ranges_batched_prepare(min_buf, max_buf, entries, reciprocal_scalar_division)()


def parameters_for_ranges_batched_prepare(min_buf: torch.Tensor, max_buf: torch.Tensor, entries,
                                          reciprocal_scalar_division: bool = False):
    """Validation and the descriptor table (host work + one small H2D copy) of ``parameters_for_ranges_batched_``
    now, the launch when the returned callable is called: the data-parallel block exit prepares before its host
    sync and launches after the ranges were all-reduced."""
    C.require_cuda(min_buf, "min_buf")
    if not entries:
        return lambda: None
    if min_buf.dtype != max_buf.dtype or not min_buf.is_contiguous() or not max_buf.is_contiguous():
        raise RuntimeError("parameters_for_ranges_batched_: range buffers must be contiguous and of one dtype")
    words = (ctypes.c_int64 * 3)()
    enc = {}
    rows = []
    for start, length, num_bits, symmetric, allow_one_sided, scale, offset in entries:
        for t in (scale, offset):
            if t is not None and (t.dtype != torch.float32 or t.numel() != length or not t.is_contiguous() or t.device != min_buf.device):
                raise RuntimeError("parameters_for_ranges_batched_: scale/offset must be contiguous fp32 tensors of the range's length")
        if start < 0 or start + length > min_buf.numel():
            raise RuntimeError("parameters_for_ranges_batched_: range outside the buffers")
        key = (float(num_bits), bool(symmetric), bool(allow_one_sided))
        if key not in enc:
            C.lib.ffq_params_for_ranges_encode(key[0], int(key[1]), _flags(key[2], reciprocal_scalar_division), words)
            enc[key] = (int(words[0]), int(words[1]), int(words[2]))
        w = enc[key]
        rows.append((int(start), int(length), scale.data_ptr(), 0 if offset is None else offset.data_ptr(), w[0], w[1], w[2], 0))
    desc = torch.tensor(rows, dtype=torch.int64).pin_memory().to(min_buf.device, non_blocking=True)
    keep = [t for e in entries for t in (e[5], e[6]) if t is not None]      # the table holds raw pointers

    def launch() -> None:
        with (C.device_of(min_buf.device) if min_buf.is_cuda else C._NO_GUARD):
            C.check(C.lib.ffq_params_for_ranges_batched(min_buf.data_ptr(), max_buf.data_ptr(), C.dtype_tag(min_buf.dtype),
                                                        desc.data_ptr(), len(rows), C.current_stream(min_buf.device)))
        del keep[:]

    return launch


def calibrate_quantize_mode(shape: Sequence[int], tile_size, dtype: torch.dtype) -> int:
    """0: the fused calibration step does not handle this layout, 1: per-channel rows, 2: per-tensor."""
    shape = tuple(shape)
    tile = shape if isinstance(tile_size, str) else tuple(int(t) for t in tile_size)
    if dtype not in (torch.float32, torch.float16, torch.bfloat16) or len(shape) != len(tile) or 0 in shape:
        return 0
    try:
        layout = C.make_layout(shape, tile)
    except (ValueError, NotImplementedError):
        return 0
    cache = layout.ws
    key = ("calq", dtype)
    if key not in cache:
        cache[key] = int(C.lib.ffq_calibrate_quantize_mode(layout.ref, C.dtype_tag(dtype)))
    return cache[key]


_CALQ_WS = int(C.lib.ffq_calibrate_quantize_workspace_bytes())


@_on_device
def calibrate_quantize_(
    run_min: torch.Tensor, run_max: torch.Tensor, data: torch.Tensor, tile_size, num_bits: float,
    symmetric: bool, allow_one_sided: bool, scale_out: torch.Tensor, offset_out: Optional[torch.Tensor],
    flags: Optional[torch.Tensor] = None, settled: Optional[torch.Tensor] = None, rowsum: bool = False,
    run_fixup: bool = True, workspace: Optional[torch.Tensor] = None, reciprocal_scalar_division: bool = False,
    stream: Optional["torch.cuda.Stream"] = None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """One RunningMinMax calibration step in one pass over ``data``: updates ``run_min``/``run_max`` in place
    (range_setting/minmax.py:229-237), writes the quantizer's ``scale``/``offset`` for the updated range in place
    (nn/linear_quantizer.py:347-357) and returns ``(int8 codes, rowsum-or-None)`` -- the codes are what
    ``quantize_by_tile(data, scale, tile, num_bits, torch.int8, offset)`` returns for those parameters, and
    ``rowsum`` (int32, one value per row of the last dimension) is what the W8A8 linear needs from them.
    ``settled`` / ``run_fixup``: see include/ffq_b200.h -- only a caller that has read ``settled != 0`` may pass
    ``run_fixup=False``.  ``workspace``: zero-initialised uint8 buffer of ``_CALQ_WS`` bytes for the per-tensor
    kernel's grid barrier, owned by the caller and used by one stream at a time (default: a cached per-(device,
    stream) buffer).  ``stream``: launch on this stream instead of the current one; the outputs are still allocated
    on the CURRENT stream, so the caller orders ``stream`` after the current stream before the call and the current
    stream after ``stream`` before the outputs are used (the estimator's overlapped parameter steps).
    Raises NotImplementedError for layouts the fused kernels do not cover (see calibrate_quantize_mode)."""
    x, shape, tile, layout = _prep(data, tile_size)
    _bitwidth_guard(torch.int8, num_bits)
    if x.numel() == 0:
        raise NotImplementedError("calibrate_quantize_: empty tensor")
    nt = layout.num_tiles
    for t, name in ((run_min, "run_min"), (run_max, "run_max")):
        C.require_cuda(t, name)
        if t.numel() != nt or not t.is_contiguous() or t.dtype != run_min.dtype or \
                torch.promote_types(t.dtype, x.dtype) != t.dtype:
            raise RuntimeError(f"{name} must be a contiguous tensor with {nt} elements whose dtype holds {x.dtype}")
    for t, name in ((scale_out, "scale_out"), (offset_out, "offset_out")):
        if t is not None and (t.dtype != torch.float32 or t.numel() != nt or not t.is_contiguous() or not t.is_cuda):
            raise RuntimeError(f"{name} must be a contiguous CUDA float32 tensor with {nt} elements")
    q = torch.empty(shape, dtype=torch.int8, device=x.device)
    row_len = shape[-1]
    rs = torch.empty(x.numel() // row_len, dtype=torch.int32, device=x.device) if rowsum else None
    ws = workspace if workspace is not None else C.barrier_workspace(x.device, _CALQ_WS)
    if ws.numel() < _CALQ_WS or ws.dtype != torch.uint8 or ws.device != x.device:
        raise RuntimeError(f"calibrate_quantize_: workspace must be a uint8 tensor of {_CALQ_WS} bytes on {x.device}")
    C.check(C.lib.ffq_calibrate_quantize(
        x.data_ptr(), C.dtype_tag(x.dtype), q.data_ptr(), run_min.data_ptr(), run_max.data_ptr(), C.dtype_tag(run_min.dtype),
        scale_out.data_ptr(), C.ptr(offset_out), C.ptr(rs), row_len, C.ptr(flags), C.ptr(settled), int(bool(run_fixup)),
        layout.ref, float(num_bits), int(bool(symmetric)), _flags(allow_one_sided, reciprocal_scalar_division),
        ws.data_ptr(), ws.numel(), C.current_stream(x.device) if stream is None else stream.cuda_stream))
    return q, rs


# ------------------------------------------------------------------------------------------
# fused MSE grid search (range_setting/min_error.py:206-216)
# ------------------------------------------------------------------------------------------
@_on_device
def grid_mse(
    data: torch.Tensor, cand_scale: torch.Tensor, cand_offset: Optional[torch.Tensor], tile_size, num_bits: float,
    quantized_dtype: Optional[torch.dtype] = None,
) -> torch.Tensor:
    """``err[c, t] = mean_tile_t((dequantize_c(quantize_c(data)) - data) ** 2)`` for every candidate
    parameter set ``(cand_scale[c], cand_offset[c])`` in ONE read of ``data`` -- the loop body of
    ``_MinAvgErrorGridEstimator.estimate_step`` with ``mse_error``.  Returns fp32 ``[C, num_tiles]``.
    Raises ``NotImplementedError`` for layouts whose tiles are not contiguous runs (the caller then
    evaluates candidate by candidate)."""
    x, shape, tile, layout = _prep(data, tile_size)
    _bitwidth_guard(quantized_dtype or x.dtype, num_bits)
    if x.numel() == 0:
        raise ValueError("grid_mse: empty tensor")
    nt = layout.num_tiles
    C.require_cuda(cand_scale, "cand_scale")
    ncand = cand_scale.shape[0]
    if cand_scale.dtype != torch.float32 or cand_scale.shape != (ncand, nt) or not cand_scale.is_contiguous():
        raise RuntimeError(f"cand_scale must be a contiguous float32 [C, {nt}] tensor")
    if cand_offset is not None and (cand_offset.dtype != torch.float32 or cand_offset.shape != cand_scale.shape
                                    or not cand_offset.is_contiguous()):
        raise RuntimeError("cand_offset must match cand_scale")
    err = torch.zeros((ncand, nt), dtype=torch.float32, device=x.device)
    ws_bytes = int(C.lib.ffq_grid_mse_workspace_bytes(layout.ref, C.dtype_tag(x.dtype), ncand))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
    C.check(C.lib.ffq_grid_mse(
        x.data_ptr(), C.dtype_tag(x.dtype), cand_scale.data_ptr(), C.ptr(cand_offset), ncand, err.data_ptr(),
        layout.ref, float(num_bits), C.ptr(ws), ws_bytes, C.current_stream(x.device)))
    return err


# ------------------------------------------------------------------------------------------
# LPBQ scale compression (export/_lpbq.py:131-160)
# ------------------------------------------------------------------------------------------
@_on_device
def lpbq_encode(scale_2d: torch.Tensor, channel_axis: int, bitwidth: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """``LPBQProcessor.grouped_dynamic_quantize`` in one kernel: for every channel of the 2-D per-block scale tensor
    (``channel_axis`` 0: a channel is a row, 1: a column) ``float_scale = max(scale) / 2**bitwidth`` and
    ``int_scale = clamp(round(scale / float_scale), 1, 2**bitwidth)``.  Returns ``(int_scale int32 like scale_2d,
    float_scale fp32 [channels])``, bit-identical to the reference's aten chain."""
    C.require_cuda(scale_2d, "scale_2d")
    if scale_2d.dim() != 2 or scale_2d.dtype != torch.float32:
        raise NotImplementedError("lpbq_encode: a 2-D float32 scale tensor is required")
    if scale_2d.numel() == 0:
        raise ValueError("lpbq_encode: empty scale tensor")
    s = scale_2d if scale_2d.is_contiguous() else scale_2d.contiguous()
    rows, cols = s.shape
    iq = torch.empty((rows, cols), dtype=torch.int32, device=s.device)
    fs = torch.empty(rows if channel_axis == 0 else cols, dtype=torch.float32, device=s.device)
    C.check(C.lib.ffq_lpbq_encode(s.data_ptr(), rows, cols, int(channel_axis), int(bitwidth), iq.data_ptr(), fs.data_ptr(),
                                  C.current_stream(s.device)))
    return iq, fs
