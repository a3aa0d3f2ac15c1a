"""CPU oracle for the GPTQ path  --  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/src/fastforward/quantization/gptq.py with the same aten ops in the same order on CPU
tensors (quantize / dequantize / parameters_for_range come from oracle/ref_ops.py).  It never imports
``fastforward``.  Parity PINNED: ``oracle/make_golden.py gptq`` runs the unmodified reference's ``gptq()`` on seeded
layers and ``tests/test_oracle_golden.py::test_gptq`` requires this file to reproduce its final weights and
parameters bit for bit.  Only tests/ may import this module.
"""

from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch

from . import ref_ops as R

Tensor = torch.Tensor


def calculate_hessian(in_features: int, activations: Sequence[Tensor]) -> Tensor:          # gptq.py:280-313
    hessian = torch.zeros((in_features, in_features), dtype=torch.float64)
    n_samples = 0
    for activation in activations:
        x = activation.to(torch.float32)
        bsz, seq_len, hidden = x.shape
        x = x.reshape(bsz * seq_len, hidden).transpose(0, 1).clone()
        hessian.mul_(n_samples / (n_samples + x.shape[1]))
        n_samples += x.shape[1]
        x.mul_(math.sqrt(2.0 / n_samples))
        hessian.add_(x @ x.transpose(0, 1))
    dead = torch.diag(hessian) == 0
    hessian[dead, dead] = 1
    return hessian.float()


def invert_hessian(hessian: Tensor, perc_damp: float) -> Tensor:                            # gptq.py:363-381
    hessian = hessian.clone()
    dampening = perc_damp * torch.mean(torch.diag(hessian))
    diag = torch.arange(hessian.shape[0])
    hessian[diag, diag] += dampening
    hessian = torch.linalg.cholesky(hessian)
    hessian = torch.cholesky_inverse(hessian)
    return torch.linalg.cholesky(hessian, upper=True)


def column_params(scale: Tensor, offset: Optional[Tensor], shape, tile, col: int):         # gptq.py:149-222
    """One (scale, offset) per row for original column `col`; parameters are flat, row-major over the tile grid."""
    rows, cols = shape
    rb, cb = tile
    view = (rows // rb, cols // cb)
    s = scale.reshape(view)[:, col // cb].repeat_interleave(rb)
    o = None if offset is None else offset.reshape(view)[:, col // cb].repeat_interleave(rb)
    return s, o


def quant_dequant_column(col: Tensor, s: Tensor, o: Optional[Tensor], num_bits: float, code_dtype) -> Tensor:   # :224-235
    q = R.quantize_by_tile(col.unsqueeze(1), s, (1, 1), num_bits, code_dtype or col.dtype, o)
    return R.dequantize_by_tile(q, s, (1, 1), o, col.dtype).flatten()


def gptq_block(block: Tensor, hinv_block: Tensor, scale: Tensor, offset: Optional[Tensor], shape, tile,
               orig_cols: Sequence[int], num_bits: float, code_dtype) -> Tuple[Tensor, Tensor, Tensor]:
    """gptq.py:106-130 for one block: returns (updated block, quantized columns, errors)."""
    wb = block.clone()
    q_out = torch.zeros_like(wb)
    err = torch.zeros_like(wb)
    for j in range(wb.shape[1]):
        s, o = column_params(scale, offset, shape, tile, int(orig_cols[j]))
        q_out[:, j] = quant_dequant_column(wb[:, j], s, o, num_bits, code_dtype)
        err[:, j] = (wb[:, j] - q_out[:, j]) / hinv_block[j, j]
        wb[:, j + 1:] -= err[:, j].unsqueeze(1) @ hinv_block[j:j + 1, j + 1:]
    return wb, q_out, err


def gptq(weight: Tensor, activations: Sequence[Tensor], scale: Tensor, offset: Optional[Tensor], tile, num_bits: float,
         symmetric: bool, allow_one_sided: bool, code_dtype=None, block_size: int = 128, perc_damp: float = 0.01,
         actorder: bool = False, grouped: bool = False):
    """gptq.py:25-147 after the initial range estimation.  `scale`/`offset` are the calibrated flat parameters
    (they are updated in place when group scales are recomputed).  Returns (new weight, errors)."""
    weights = weight.clone().float()
    rows, columns = weights.shape
    hessian = calculate_hessian(columns, activations)
    order = torch.argsort(torch.diag(hessian), descending=True) if actorder else torch.arange(columns)
    weights = weights[:, order]
    hessian = hessian[order][:, order]
    quantized = torch.zeros_like(weights)
    errors = torch.zeros_like(weights)
    hinv = invert_hessian(hessian, perc_damp)
    rb, cb = tile
    nrb, ncb = rows // rb, columns // cb
    recompute = grouped and ncb > 1 and not actorder
    for i in range(0, columns, block_size):
        end = min(i + block_size, columns)
        wb = weights[:, i:end].clone()
        hb = hinv[i:end, i:end]
        for j in range(end - i):
            gc = i + j
            if recompute and gc % cb == 0:
                reshaped = weights[:, gc:gc + cb].reshape(nrb, -1)
                s_new, o_new = R.parameters_for_range(reshaped.min(dim=-1).values, reshaped.max(dim=-1).values, num_bits,
                                                      symmetric, allow_one_sided)
                scale.view(nrb, ncb)[:, gc // cb] = s_new.to(scale.dtype)
                if offset is not None:
                    offset.view(nrb, ncb)[:, gc // cb] = 0.0 if o_new is None else o_new.to(offset.dtype)
            s, o = column_params(scale, offset, (rows, columns), tile, int(order[gc]))
            quantized[:, gc] = quant_dequant_column(wb[:, j], s, o, num_bits, code_dtype)
            errors[:, gc] = (wb[:, j] - quantized[:, gc]) / hb[j, j]
            wb[:, j + 1:] -= errors[:, gc].unsqueeze(1) @ hb[j:j + 1, j + 1:]
        weights[:, end:] -= errors[:, i:end] @ hinv[i:end, end:]
    restore = torch.argsort(order)
    return quantized[:, restore].to(weight.dtype), errors[:, restore]
