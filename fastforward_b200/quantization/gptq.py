"""GPTQ for ``QuantizedLinear`` (reference: quantization/gptq.py:25-381; same function names, arguments and errors).

What changed for B200: the per-column inner loop (gptq.py:106-130 -- quantize, dequantize, error, rank-1 update of
the rest of the block: ~15 kernel launches per COLUMN in the reference) is ONE kernel per column block
(``ffq_gptq_block``: a warp owns a row of the block in registers, the Hinv block sits in shared memory), with the
reference's op-by-op rounding, so a block's outputs are bit-identical to the reference's for the same inputs.  The
Hessian accumulation, the Cholesky inversion and the trailing update between blocks are plain library calls
(GEMM / cuSOLVER through torch), as in the reference.
"""

from __future__ import annotations

import sys

import logging
import math
from typing import Any, Callable, Iterable, Optional

import torch

from .. import _cabi as C
from .. import ops
from . import affine as affine_quant
from . import granularity as granularities

logger = logging.getLogger(__name__)

_MAX_BLOCK = 128      # columns ffq_gptq_block keeps in registers


def _check_granularity(granularity) -> None:
    if isinstance(granularity, granularities.PerBlock) and not granularity.strict_blocks:
        raise ValueError("GPTQ does not support PerBlock with strict_blocks=False.")
    ok = isinstance(granularity, (granularities.PerTensor, granularities.PerBlock, granularities.PerTile)) or (
        isinstance(granularity, granularities.PerChannel) and tuple(granularity.channel_dims) in ((0,), (1,), (0, 1)))
    if not ok:
        raise TypeError(f"Unsupported granularity: {type(granularity).__name__}")


def _tile(granularity, shape) -> tuple:
    tile = granularity.tile_size(shape)
    return tuple(shape) if isinstance(tile, str) else tuple(int(t) for t in tile)


def gptq(module, dataset: Iterable, block_size: int = 128, perc_damp: float = 0.01, actorder: bool = False,
         layer_name: str = "") -> None:
    """Quantize a ``QuantizedLinear`` in place with GPTQ (gptq.py:25-147)."""
    from ..nn.linear import QuantizedLinear
    from ..nn.linear_quantizer import LinearQuantizer
    from ..range_setting import estimate_ranges, smoothed_minmax

    if not isinstance(module.weight_quantizer, LinearQuantizer):
        raise ValueError(f"weight_quantizer must be a LinearQuantizer, got {type(module.weight_quantizer).__name__}.")
    if not isinstance(module, QuantizedLinear) or module.weight.dim() != 2:
        raise NotImplementedError("fastforward_b200.gptq supports QuantizedLinear layers")
    weight_quantizer = module.weight_quantizer
    _check_granularity(weight_quantizer.granularity)
    if block_size < 1 or block_size > _MAX_BLOCK:
        raise NotImplementedError(f"fastforward_b200.gptq: block_size must be in [1, {_MAX_BLOCK}]")
    C.require_cuda(module.weight, "module.weight")

    original_shape = module.weight.shape
    weights = module.weight.data.clone().float()
    rows, columns = weights.shape

    with torch.no_grad(), estimate_ranges(weight_quantizer, smoothed_minmax):
        weight_quantizer(weights)

    hessian = calculate_hessian(module, dataset)
    column_order = torch.argsort(torch.diag(hessian), descending=True) if actorder else \
        torch.arange(columns, device=weights.device)
    weights = weights[:, column_order].contiguous()
    hessian = hessian[column_order][:, column_order]

    quantized_weights = torch.zeros_like(weights)
    errors = torch.zeros_like(weights)
    hessian_inverse = invert_hessian(hessian, perc_damp).contiguous()

    tile = _tile(weight_quantizer.granularity, weights.shape)
    row_block, col_block = tile
    num_row_blocks, num_col_blocks = rows // row_block, columns // col_block
    # grouped quantization only: each group's scale is recomputed on its error-corrected weights (gptq.py:91-100)
    recompute = isinstance(weight_quantizer.granularity, (granularities.PerBlock, granularities.PerTile)) and \
        num_col_blocks > 1 and not actorder
    order32 = column_order.to(torch.int32).contiguous()

    with torch.no_grad():
        for i in range(0, columns, block_size):
            end = min(i + block_size, columns)
            if recompute:
                # the reference reads `weights` (not the in-block corrected copy) whenever a group starts inside
                # this block, and `weights[:, i:end]` does not change during the block: all those ranges are known now
                first_group = -(-i // col_block) * col_block
                for gc in range(first_group, end, col_block):
                    group = weights[:, gc:gc + col_block].contiguous()
                    lo, hi = ops.tile_minmax(group, (row_block, group.shape[1]))
                    update_partial_range(weight_quantizer, lo, hi, param_view_shape=(num_row_blocks, num_col_blocks),
                                         param_view_index=(slice(None), gc // col_block))
            block = weights[:, i:end].contiguous()
            gptq_block_(block, quantized_weights[:, i:end], errors[:, i:end], hessian_inverse[i:end, i:end],
                        weight_quantizer.scale.data, None if weight_quantizer.offset is None else weight_quantizer.offset.data,
                        order32[i:end], row_block, col_block, num_col_blocks, weight_quantizer.num_bits,
                        weight_quantizer.quantized_dtype)
            if end < columns:
                weights[:, end:] -= errors[:, i:end] @ hessian_inverse[i:end, end:]

        restore = torch.argsort(column_order)
        quantized_weights = quantized_weights[:, restore]
        errors = errors[:, restore]
        module.weight.copy_(quantized_weights.view(original_shape).to(module.weight.dtype))
    loss = torch.mean(torch.abs(errors)).item()
    logger.info("[GPTQ][wbits=%d][%s] loss=%.6f", weight_quantizer.num_bits, layer_name, loss)


def gptq_block_(block: torch.Tensor, q_out: torch.Tensor, err_out: torch.Tensor, hinv_block: torch.Tensor,
                scale: torch.Tensor, offset: Optional[torch.Tensor], orig_col: torch.Tensor, row_block: int, col_block: int,
                num_col_blocks: int, num_bits: float, quantized_dtype: Optional[torch.dtype] = None) -> None:
    """The loop over the columns of one block (gptq.py:106-130) as one kernel.  ``block`` ([R, n] fp32, row stride
    arbitrary) is updated in place like the reference's ``weights_block``; ``q_out`` / ``err_out`` ([R, n] fp32 views,
    e.g. column slices of the full matrices) receive the quantize-dequantized columns and the errors; ``hinv_block``
    is ``Hinv[i:i+n, i:i+n]``; ``orig_col`` (int32 [n]) holds each column's index in the un-permuted weight."""
    for t, name in ((block, "block"), (q_out, "q_out"), (err_out, "err_out"), (hinv_block, "hinv_block")):
        C.require_cuda(t, name)
        if t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
            raise RuntimeError(f"gptq_block_: '{name}' must be a 2-D float32 tensor with unit column stride")
    r, n = block.shape
    if q_out.shape != block.shape or err_out.shape != block.shape or hinv_block.shape != (n, n) or orig_col.numel() != n:
        raise RuntimeError("gptq_block_: shape mismatch")
    if orig_col.dtype != torch.int32 or not orig_col.is_contiguous():
        raise RuntimeError("gptq_block_: orig_col must be a contiguous int32 tensor")
    need = (r // row_block) * num_col_blocks if row_block else 0
    s = scale.detach().reshape(-1)
    o = None if offset is None else offset.detach().reshape(-1)
    if s.numel() != need or (o is not None and o.numel() != need) or not s.is_contiguous():
        raise RuntimeError(f"gptq_block_: expected {need} quantization parameters, got {s.numel()}")
    code_dtype = quantized_dtype or torch.float32
    ops._bitwidth_guard(code_dtype, num_bits)
    C.check(C.lib.ffq_gptq_block(
        block.data_ptr(), block.stride(0), q_out.data_ptr(), q_out.stride(0), err_out.data_ptr(), err_out.stride(0),
        hinv_block.data_ptr(), hinv_block.stride(0), r, n, s.data_ptr(), C.dtype_tag(s.dtype), C.ptr(o),
        C.dtype_tag(o.dtype if o is not None else None), orig_col.data_ptr(), row_block, col_block, num_col_blocks,
        float(num_bits), C.dtype_tag(code_dtype), C.current_stream(block.device)))


def column_quantizer(weight_quantizer, weight_shape, col_index: int) -> Callable[[torch.Tensor], torch.Tensor]:
    """Quantize-dequantize operator for one column with one (scale, offset) per row (gptq.py:149-235)."""
    out_features, in_features = weight_shape
    granularity = weight_quantizer.granularity
    _check_granularity(granularity)
    row_block, col_block = _tile(granularity, weight_shape)
    num_col_blocks = in_features // col_block
    view = (out_features // row_block, num_col_blocks)
    scale = weight_quantizer.scale.detach().reshape(view)[:, col_index // col_block].repeat_interleave(row_block)
    offset = None if weight_quantizer.offset is None else \
        weight_quantizer.offset.detach().reshape(view)[:, col_index // col_block].repeat_interleave(row_block)
    ctx = affine_quant.quantization_context(scale=scale.contiguous(), offset=None if offset is None else offset.contiguous(),
                                            num_bits=weight_quantizer.num_bits, granularity=granularities.PerChannel(0),
                                            output_dtype=weight_quantizer.quantized_dtype)

    def _quant_fn(col: torch.Tensor) -> torch.Tensor:
        q = ctx.quantization_fn.quantize(col.unsqueeze(1), ctx.quantization_params)
        return q.dequantize().flatten()

    return _quant_fn


def update_partial_range(weight_quantizer, min_range: torch.Tensor, max_range: torch.Tensor, *, param_view_shape,
                         param_view_index: Any) -> None:
    """Write scale/offset for the selected parameter positions from a min/max range (gptq.py:238-277)."""
    scale, offset = affine_quant.parameters_for_range(
        min_range, max_range, num_bits=weight_quantizer.num_bits, symmetric=weight_quantizer.symmetric,
        allow_one_sided=weight_quantizer.allow_one_sided)
    scale_view = weight_quantizer.scale.data.view(param_view_shape)
    scale_view[param_view_index] = scale.to(scale_view.dtype)
    if weight_quantizer.offset is not None:
        offset_view = weight_quantizer.offset.data.view(param_view_shape)
        if offset is not None:
            offset_view[param_view_index] = offset.to(offset_view.dtype)
        else:
            offset_view[param_view_index] = 0.0


def calculate_hessian(layer, activations: Iterable) -> torch.Tensor:
    """Running average of ``2 X X^T`` over the calibration inputs, float64 accumulation (gptq.py:280-313)."""
    device = layer.weight.device
    in_features = layer.weight.shape[1]
    hessian = torch.zeros((in_features, in_features), device=device, dtype=torch.float64)
    n_samples = 0
    for (activation,), _ in activations:
        x = activation.to(device=device, dtype=torch.float32)
        x = x.reshape(-1, x.shape[-1]).transpose(0, 1)        # (B, S, H) -> (H, B*S)
        hessian.mul_(n_samples / (n_samples + x.shape[1]))
        n_samples += x.shape[1]
        x = x * math.sqrt(2.0 / n_samples)
        hessian.add_(x @ x.transpose(0, 1))
    dead = torch.diag(hessian) == 0
    hessian[dead, dead] = 1
    return hessian.float()


def invert_hessian(hessian: torch.Tensor, perc_damp: float) -> torch.Tensor:
    """Upper Cholesky factor of the damped inverse Hessian (gptq.py:363-381)."""
    dampening = perc_damp * torch.mean(torch.diag(hessian))
    diag = torch.arange(hessian.shape[0], device=hessian.device)
    hessian[diag, diag] += dampening
    hessian = torch.linalg.cholesky(hessian)
    hessian = torch.cholesky_inverse(hessian)
    hessian = torch.linalg.cholesky(hessian, upper=True)
    return hessian


# ``ff.quantization.gptq(module, dataset, ...)`` is a function in the reference (quantization/__init__.py:15) and a
# submodule here (its helpers -- gptq_block_ and friends -- are imported from it): the module itself is callable.
class _CallableModule(type(sys)):
    def __call__(self, *args, **kwargs):
        return self.gptq(*args, **kwargs)


sys.modules[__name__].__class__ = _CallableModule
