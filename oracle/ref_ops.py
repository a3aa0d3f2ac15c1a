"""CPU oracle for the FastForward quantization hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement of the reference's algorithm for the hot path
(SURVEY.md section 8a).  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``fastforward_b200/`` imports it; the product path
has no CPU fallback and fails loudly when the CUDA library is missing.

Why torch-on-CPU and not numpy/C: the reference has no arithmetic of its own -- its
four ops are chains of PyTorch aten eager ops (third-party dependency ``torch>=2.4``,
installed 2.11.0).  Bit-exactness is therefore defined by aten's IEEE semantics
(true division, round-half-even ``round``, NaN-propagating ``clamp``/``min``/``max``,
per-op rounding to the promoted dtype).  The restatement below issues the same aten
ops in the same order, on CPU tensors, so it inherits exactly those semantics, and
is independent of the reference's Python (it never imports ``fastforward``).  A
plain-C scalar restatement of the fp32 path lives beside it (``ffq_oracle.c``) as an
independent cross-check of the IEEE claims.

Parity pinning: ``oracle/make_golden.py`` runs the UNMODIFIED reference (imported from
/root/reference/src through a two-file shim) on seeded inputs and stores inputs and
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file
against every stored vector (and against the reference's own hand-written golden
vectors, e.g. tests/nn/test_linear_quantizer.py:20-59).  Status: parity PINNED for
codes, dequantized values, ranges, scale/offset and dx; per-tile gradient *sums* are
pinned only to a tolerance, by construction (SURVEY.md section 7 'hard parts').

All ``file:line`` citations are into /root/reference/src/fastforward/.
"""

from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# tile <-> row layout                                    quantization/tiled_tensor.py:71-144
# --------------------------------------------------------------------------------------
def _check_tiles(shape: Sequence[int], tile: Sequence[int]) -> None:
    # quantization/tiled_tensor.py:19-42 : same error type (ValueError) for both failures
    if len(shape) != len(tile):
        raise ValueError(
            f"Input dimensionality must match tile_size dimensionality got {len(shape)} and {len(tile)}"
        )
    bad = [i for i, (d, t) in enumerate(zip(shape, tile)) if t > 0 and d % t != 0]
    if bad:
        raise ValueError(
            "Each dimension of tile_size must divide the corresponding input dimension. Got "
            + ", ".join(f"{shape[i]} and {tile[i]} for dimension {i}" for i in bad)
            + "."
        )


def tile_rows(data: Tensor, tile: Sequence[int]) -> Tensor:
    """[d0..dn] -> [num_tiles, tile_numel]; tile index row-major over the block grid,
    within-tile order row-major over the tile (tiled_tensor.py:71-98)."""
    if data.numel() == 0:
        return data.reshape(1, 0)
    tile = tuple(int(t) for t in tile)
    _check_tiles(tuple(data.shape), tile)
    split = []
    for d, t in zip(data.shape, tile):
        split += [d // t, t]
    n = len(tile)
    order = [2 * i for i in range(n)] + [2 * i + 1 for i in range(n)]
    ntiles = data.numel() // max(1, math.prod(tile))
    return data.reshape(split).permute(order).reshape(ntiles, -1)


def untile_rows(rows: Tensor, shape: Sequence[int], tile: Sequence[int]) -> Tensor:
    """Inverse of :func:`tile_rows` (tiled_tensor.py:101-144)."""
    shape = tuple(int(s) for s in shape)
    if rows.numel() == 0:
        return rows.reshape(shape)
    tile = tuple(int(t) for t in tile)
    _check_tiles(shape, tile)
    n = len(tile)
    grid = [d // t for d, t in zip(shape, tile)]
    back = []
    for i in range(n):
        back += [i, n + i]
    return rows.reshape(grid + list(tile)).permute(back).reshape(shape)


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def int_min(num_bits: float) -> float:  # quantization/affine/range.py:9-17
    return -(2 ** (num_bits - 1))


def int_max(num_bits: float) -> float:  # quantization/affine/range.py:20-28
    return -int_min(num_bits) - 1


_MANTISSA = {
    torch.bfloat16: 7,
    torch.float16: 10,
    torch.float32: 23,
    torch.float64: 52,
}


def can_support_bitwidth(dtype: torch.dtype, num_bits: float) -> bool:
    """quantization/_quantizer_impl.py:44-75 (mantissa + 2 >= bits; ints: iinfo.bits + 2)."""
    if dtype.is_floating_point:
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


def _rounded_offset(offset: Optional[Tensor], scale: Tensor) -> Tensor:
    # _quantizer_impl.py:140-141
    if offset is None:
        return torch.zeros_like(scale)
    return torch.round(offset.reshape(-1))


# --------------------------------------------------------------------------------------
# a1  quantize_by_tile                               quantization/_quantizer_impl.py:144-169
# --------------------------------------------------------------------------------------
def quantize_by_tile(
    data: Tensor,
    scale: Tensor,
    tile: Sequence[int],
    num_bits: float,
    output_dtype: Optional[torch.dtype],
    offset: Optional[Tensor] = None,
) -> Tensor:
    s = scale.reshape(-1)
    o = _rounded_offset(offset, s)
    lo, hi = int_min(num_bits), int_max(num_bits)
    rows = tile_rows(data, tile)
    q = torch.round(rows / s[:, None] - o[:, None])          # :161
    q = torch.clamp(q, lo, hi)                               # :162
    out = untile_rows(q, data.shape, tile)
    output_dtype = output_dtype or out.dtype
    if not can_support_bitwidth(output_dtype, num_bits):     # :165-167
        raise RuntimeError(
            f"Provided dtype ({output_dtype}) is not enough to store {num_bits} bits quantized values."
        )
    return out.to(output_dtype)


# --------------------------------------------------------------------------------------
# a2  dequantize_by_tile                             quantization/_quantizer_impl.py:172-190
# --------------------------------------------------------------------------------------
def dequantize_by_tile(
    data: Tensor,
    scale: Tensor,
    tile: Sequence[int],
    offset: Optional[Tensor] = None,
    output_dtype: Optional[torch.dtype] = None,
) -> Tensor:
    s = scale.reshape(-1)
    o = _rounded_offset(offset, s)
    rows = tile_rows(data, tile)
    y = (rows + o[:, None]) * s[:, None]                     # :186
    y = untile_rows(y, data.shape, tile)
    if output_dtype:
        y = y.to(output_dtype)
    return y


# --------------------------------------------------------------------------------------
# a3  quantize_by_tile_backward (STE)                quantization/_quantizer_impl.py:193-237
# --------------------------------------------------------------------------------------
def backward_terms(
    data: Tensor,
    grad: Tensor,
    scale: Tensor,
    tile: Sequence[int],
    num_bits: float,
    offset: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
    """Return (dx, dscale_elem_rows, doffset_elem_rows): the *pre-reduction* per-element
    terms in row layout -- these are bit-exact quantities; the row sums are not."""
    s = scale.reshape(-1)
    o = _rounded_offset(offset, s)
    lo, hi = int_min(num_bits), int_max(num_bits)
    xr = tile_rows(data, tile)
    gr = tile_rows(grad, tile)
    pre = (xr / s[:, None]) - o[:, None]                     # :212
    q = torch.round(pre)                                     # :213
    clip = torch.logical_or(q < lo, q > hi)                  # :214
    dx = untile_rows(torch.where(clip, 0, gr), data.shape, tile)      # :216, :235
    doff = None
    if offset is not None:
        doff = torch.where(clip, s[:, None] * gr, 0)         # :221
    dsc = torch.empty(q.shape, dtype=s.dtype, device=s.device)   # :224
    torch.where(q < lo, s.new_tensor([lo]), s.new_tensor([hi]), out=dsc)   # :225-227
    dsc.add_(o[:, None].to(dsc.dtype))                       # :228
    torch.where(clip, dsc, (q - pre).to(dsc.dtype), out=dsc)  # :229
    dsc.mul_(gr)                                             # :230
    return dx, dsc, doff


def quantize_by_tile_backward(
    data: Tensor,
    grad: Tensor,
    scale: Tensor,
    tile: Sequence[int],
    num_bits: float,
    offset: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Tensor]:
    """(dx, dscale, doffset) exactly as the reference sums them (aten ``sum(1)``)."""
    dx, dsc, doff = backward_terms(data, grad, scale, tile, num_bits, offset)
    dscale = dsc.sum(1).reshape(scale.shape)                 # :236
    doffset = torch.Tensor() if doff is None else doff.sum(1).reshape(scale.shape)  # :219-222
    return dx, dscale, doffset


def quantize_by_tile_backward_f64(
    data: Tensor, grad: Tensor, scale: Tensor, tile: Sequence[int], num_bits: float,
    offset: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
    """Same, but the per-tile sums are accumulated in float64 from the bit-exact
    per-element terms: the order-independent target both sides are compared to."""
    dx, dsc, doff = backward_terms(data, grad, scale, tile, num_bits, offset)
    dscale = dsc.double().sum(1).reshape(scale.shape)
    doffset = None if doff is None else doff.double().sum(1).reshape(scale.shape)
    return dx, dscale, doffset


# --------------------------------------------------------------------------------------
# a6  parameters_for_range / quantization_range      quantization/affine/range.py:31-122
# --------------------------------------------------------------------------------------
def parameters_for_range(
    min_range: Tensor, max_range: Tensor, num_bits: float, symmetric: bool, allow_one_sided: bool
) -> Tuple[Tensor, Optional[Tensor]]:
    mn = torch.as_tensor(min_range).to(torch.float32)        # :89-90
    mx = torch.as_tensor(max_range).to(torch.float32)
    one_sided = bool(mn.min() >= 0) and allow_one_sided      # :100 (global over all tiles)
    lo = int_min(num_bits)
    if symmetric and one_sided:
        mn = torch.zeros_like(mn)                            # :104-105
    if symmetric and not one_sided:
        neg = torch.abs(mn) / abs(lo)                        # :109
        pos = torch.abs(mx) / abs(int_max(num_bits))         # :110
        return torch.max(neg, pos), None
    steps = 2 ** num_bits - 1                                # :117
    sc = (mx - mn) / steps
    sc = sc.clamp(torch.finfo(sc.dtype).eps)                 # :120
    return sc, mn / sc - lo                                  # :121


def quantization_range(scale, offset, num_bits: float):
    off = 0.0 if offset is None else offset                  # range.py:48-51
    return (int_min(num_bits) + off) * scale, (int_max(num_bits) + off) * scale


# --------------------------------------------------------------------------------------
# a4  quantize_dynamic_by_tile                       quantization/_quantizer_impl.py:243-285
# --------------------------------------------------------------------------------------
def quantize_dynamic_by_tile(
    data: Tensor, tile: Sequence[int], num_bits: float, symmetric: bool, allow_one_sided: bool,
    output_dtype: Optional[torch.dtype],
) -> Tuple[Tensor, Tensor, Tensor]:
    lo, hi = int_min(num_bits), int_max(num_bits)
    rows = tile_rows(data, tile)
    if rows.numel() == 0:                                    # :259-264 (QuantizationError there)
        raise ValueError(f"Cannot dynamically quantize an empty tensor of shape {data.shape}")
    mn = torch.min(rows, dim=1).values
    mx = torch.max(rows, dim=1).values
    scale, offset = parameters_for_range(mn, mx, num_bits, symmetric, allow_one_sided)
    if offset is None:
        offset = torch.zeros_like(scale)
    offset = torch.round(offset)                             # :275
    q = torch.round(rows / scale[:, None] - offset[:, None])
    q = torch.clamp(q, lo, hi)
    out = untile_rows(q, data.shape, tile)
    output_dtype = output_dtype or out.dtype
    if not can_support_bitwidth(output_dtype, num_bits):
        raise RuntimeError(
            f"Provided dtype ({output_dtype}) is not enough to store {num_bits} bits quantized values."
        )
    return out.to(output_dtype), scale, offset


# --------------------------------------------------------------------------------------
# a7  RunningMinMax step                             range_setting/minmax.py:215-239
# --------------------------------------------------------------------------------------
def tile_minmax(data: Tensor, tile: Sequence[int]) -> Tuple[Tensor, Tensor]:
    rows = tile_rows(data, tile)                             # minmax.py:227-230
    return torch.min(rows, -1).values, torch.max(rows, -1).values


def running_minmax_step(
    run_min: Optional[Tensor], run_max: Optional[Tensor], data: Tensor, tile: Sequence[int]
) -> Tuple[Tensor, Tensor]:
    mn, mx = tile_minmax(data, tile)
    if bool(mn.isinf().any()) or bool(mx.isinf().any()):     # minmax.py:233-234
        raise NotImplementedError("Infinite")
    if run_min is None:                                      # minmax.py:209-213
        run_min = data.new_full(mn.shape, float("inf"))
    if run_max is None:
        run_max = data.new_full(mx.shape, float("-inf"))
    return torch.min(run_min, mn), torch.max(run_max, mx)    # :236-237


def smoothed_minmax_step(
    run_min: Optional[Tensor], run_max: Optional[Tensor], data: Tensor, tile: Sequence[int], gamma: float
) -> Tuple[Tensor, Tensor]:
    mn, mx = tile_minmax(data, tile)                         # minmax.py:79-90
    if run_min is None or run_max is None or bool(run_min.isinf().any()) or bool(run_max.isinf().any()):
        return mn, mx
    return gamma * mn + (1 - gamma) * run_min, gamma * mx + (1 - gamma) * run_max


# --------------------------------------------------------------------------------------
# a12 quantized linear: the reference's dequantize-then-float fallback
#                                                    _gen/fallback.py:77-112
# --------------------------------------------------------------------------------------
def fallback_linear(
    x_codes: Tensor, x_scale: Tensor, x_offset: Optional[Tensor], x_tile: Sequence[int], x_dtype: torch.dtype,
    w_codes: Tensor, w_scale: Tensor, w_offset: Optional[Tensor], w_tile: Sequence[int], w_dtype: torch.dtype,
    bias: Optional[Tensor] = None,
) -> Tensor:
    x = dequantize_by_tile(x_codes, x_scale, x_tile, x_offset, x_dtype)     # fallback.py:94-95
    w = dequantize_by_tile(w_codes, w_scale, w_tile, w_offset, w_dtype)     # :102-103
    return torch.nn.functional.linear(x, w, bias)                           # :108


def exact_linear_f64(
    x_codes: Tensor, x_scale: Tensor, x_offset: Optional[Tensor], x_tile: Sequence[int],
    w_codes: Tensor, w_scale: Tensor, w_offset: Optional[Tensor], w_tile: Sequence[int],
    bias: Optional[Tensor] = None,
) -> Tensor:
    """float64 dequantized matmul -- the accuracy yardstick for both the int8 kernel and
    the float fallback."""
    x = dequantize_by_tile(x_codes.double(), x_scale.double(), x_tile,
                           None if x_offset is None else x_offset.double())
    w = dequantize_by_tile(w_codes.double(), w_scale.double(), w_tile,
                           None if w_offset is None else w_offset.double())
    return torch.nn.functional.linear(x, w, None if bias is None else bias.double())


# --------------------------------------------------------------------------------------
# composite steps used by bench.py's CPU baseline (same call sequence as the reference)
# --------------------------------------------------------------------------------------
def fake_quant_fwd_bwd(
    x: Tensor, g: Tensor, scale: Tensor, offset: Optional[Tensor], tile: Sequence[int], num_bits: float
) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """``q(x).dequantize().backward(g)``: a1 -> a2 -> (identity) -> a3
    (affine/_autograd.py:68-104,136-156)."""
    q = quantize_by_tile(x, scale, tile, num_bits, x.dtype, offset)
    y = dequantize_by_tile(q, scale, tile, offset, x.dtype)
    dx, dscale, doffset = quantize_by_tile_backward(x, g, scale, tile, num_bits, offset)
    return y, dx, dscale, doffset


def calibration_quantizer_step(
    run_min: Optional[Tensor], run_max: Optional[Tensor], data: Tensor, tile: Sequence[int],
    num_bits: float, symmetric: bool, allow_one_sided: bool, quantized_dtype: Optional[torch.dtype],
):
    """One RunningMinMax calibration step of one quantizer followed by the quantize call
    (range_setting/common.py:218-238 -> minmax.py:215-239 -> nn/linear_quantizer.py:327-357).

    Returns (run_min, run_max, scale, offset_or_None_as_stored, codes)."""
    run_min, run_max = running_minmax_step(run_min, run_max, data, tile)
    scale, offset = parameters_for_range(run_min, run_max, num_bits, symmetric, allow_one_sided)
    # stored params (linear_quantizer.py:350-357): offset buffer is 0 when the computed one is None
    if offset is None and (symmetric and allow_one_sided):
        offset = torch.zeros_like(scale)
    codes = quantize_by_tile(data, scale, tile, num_bits, quantized_dtype or data.dtype, offset)
    return run_min, run_max, scale, offset, codes


# --------------------------------------------------------------------------------------
# f2  MinErrorGridRangeEstimator (mse_grid)          range_setting/min_error.py:64-221
# --------------------------------------------------------------------------------------
def uniform_search_grid(
    data: Tensor, tile: Sequence[int], symmetric: bool, num_candidates: int,
    absolute_margin: float = 0.5, relative_margin: float = 1.0,
) -> Tuple[Tensor, Tensor]:
    """Candidate (min_threshold, max_threshold), each ``[num_candidates, num_tiles]``
    (min_error.py:78-146; three branches: non-negative data, asymmetric, symmetric)."""
    rows = tile_rows(data, tile)
    max_data = relative_margin * rows.max(dim=1).values + absolute_margin          # :102
    min_data = relative_margin * rows.min(dim=1).values - absolute_margin          # :103
    kw = dict(dtype=rows.dtype, device=rows.device)
    if not bool(min_data.min() < 0):                                               # :110-115
        lo = torch.zeros((num_candidates, rows.shape[0]), **kw)
        steps = torch.linspace(1 / num_candidates, 1, num_candidates, **kw)
        hi = steps.unsqueeze(1) * max_data.unsqueeze(0)
    elif not symmetric:                                                            # :117-139
        margin = 0.6
        n_lo = math.floor(math.sqrt(num_candidates))
        n_hi = n_lo + num_candidates - n_lo ** 2
        steps_lo = torch.linspace(1, margin, n_lo, **kw)
        steps_hi = torch.linspace(margin, 1, n_hi, **kw)
        lo = steps_lo.unsqueeze(1) * (relative_margin * min_data.unsqueeze(0) + absolute_margin)
        hi = steps_hi.unsqueeze(1) * (relative_margin * max_data.unsqueeze(0) + absolute_margin)
        lo = lo.repeat(n_hi, 1)
        hi = hi.repeat_interleave(n_lo, dim=0)
    else:                                                                          # :141-146
        steps = torch.linspace(1 / num_candidates, 1, num_candidates, **kw)
        hi = steps.unsqueeze(1) * torch.max(torch.abs(min_data), torch.abs(max_data)).unsqueeze(0)
        lo = -hi
    return lo, hi


def mse_grid_errors(
    data: Tensor, tile: Sequence[int], lo: Tensor, hi: Tensor, num_bits: float, symmetric: bool,
    allow_one_sided: bool, quantized_dtype: Optional[torch.dtype] = None, num_candidates: Optional[int] = None,
) -> Tensor:
    """Per-candidate, per-tile mean squared error of one batch, ``[len(lo), num_tiles]``
    (min_error.py:206-216: operator_for_range -> quantize -> dequantize -> mse_error :64-74).

    The reference loops ``range(self.num_candidates)`` (:207) although the asymmetric grid holds
    ``n_lo * n_hi >= num_candidates`` rows (:121-139): rows past ``num_candidates`` are never
    evaluated and keep an accumulated error of 0 -- reproduced here, it decides the argmin."""
    rows = tile_rows(data, tile)
    errs = []
    evaluated = lo.shape[0] if num_candidates is None else num_candidates
    for i in range(lo.shape[0]):
        if i >= evaluated:
            errs.append(torch.zeros(rows.shape[0], dtype=data.dtype))
            continue
        scale, offset = parameters_for_range(lo[i], hi[i], num_bits, symmetric, allow_one_sided)
        q = quantize_by_tile(data, scale, tile, num_bits, quantized_dtype or data.dtype, offset)
        y = dequantize_by_tile(q, scale, tile, offset, data.dtype)
        errs.append(torch.mean((tile_rows(y, tile) - rows) ** 2, dim=1))
    return torch.stack(errs)


def mse_grid_select(cumulative_error: Tensor, lo: Tensor, hi: Tensor) -> Tuple[Tensor, Tensor]:
    """Range of the candidate with the smallest accumulated error, per tile (min_error.py:197-204)."""
    best = cumulative_error.min(dim=0).indices
    idx = torch.arange(lo.shape[1])
    return lo[best, idx], hi[best, idx]


# --------------------------------------------------------------------------------------
# f4  LPBQ scale compression                           export/_lpbq.py:131-160
# --------------------------------------------------------------------------------------
def lpbq_grouped_dynamic_quantize(scale_2d: Tensor, channel_axis: int, bitwidth: int) -> Tuple[Tensor, Tensor]:
    """``LPBQProcessor.grouped_dynamic_quantize`` restated: per channel (``channel_axis`` 0: a row of ``scale_2d``,
    1: a column) the float scale is the channel's largest block scale divided by ``2**bitwidth`` (:149-150) and every
    block scale becomes ``clamp(round(scale / float_scale), 1, 2**bitwidth)`` (:153-155).  Returns (integers as
    int64 shaped like ``scale_2d``, float scale with the reduced dimension kept)."""
    reduce_dim = 1 if channel_axis == 0 else 0
    max_scale = torch.amax(scale_2d, dim=reduce_dim, keepdim=True)
    dynamic_scale = max_scale / max_scale.new_tensor(2 ** bitwidth)
    q = torch.clamp(torch.round(scale_2d / dynamic_scale), 1, 2 ** bitwidth)
    return q.to(torch.int64), dynamic_scale
