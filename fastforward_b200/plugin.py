"""Drop-in registration against an installed, UNMODIFIED ``fastforward`` package.

``install()`` puts the B200 kernels behind the reference's own seams:

  seam 1  ``torch.ops.fastforward.{quantize_by_tile, dequantize_by_tile, quantize_by_tile_backward,
          quantize_dynamic_by_tile}`` are ``torch.library.custom_op``s registered device-agnostically
          (quantization/_quantizer_impl.py:127-134,144,172,193,243).  We add a CUDA-key kernel to each
          with ``CustomOpDef.register_kernel("cuda")``; the dispatcher then routes CUDA tensors to the
          hand-written kernels and leaves CPU tensors on the reference's eager chain.  Nothing above
          the ops changes: LinearQuantizer, QuantizedTensor, autograd wrappers all keep working.
  seam 2  ``fastforward.dispatcher.register("linear", predicate, kernel)`` (dispatcher.py:233-265)
          receives the W8A8 tensor-core linear for int8 per-tensor x int8 per-channel operands.
  seam 3  ``fastforward.range_setting.running_minmax`` keeps its name; ``install(patch_estimators=True)``
          swaps in the sync-free estimator so ``ff.estimate_ranges(model, ff.range_setting.running_minmax)``
          uses the fused min/max kernel.

The argument facts relied on are the ones observed at the seam: ``num_bits`` arrives as a float,
``output_dtype`` is never None from the reference's callers, ``tile_size`` is already resolved.
"""

from __future__ import annotations

from typing import Any, List, Optional

import torch

from . import ops

_installed: dict = {}


def _cuda_quantize(data, scale, tile_size, num_bits, output_dtype, offset=None):
    return ops.quantize_by_tile(data, scale, tuple(tile_size), num_bits, output_dtype, offset)


def _cuda_dequantize(data, scale, tile_size, offset=None, output_dtype=None):
    return ops.dequantize_by_tile(data, scale, tuple(tile_size), offset, output_dtype)


def _cuda_backward(data, output_grad, scale, tile_size, num_bits, offset=None) -> List[torch.Tensor]:
    dx, dscale, doffset = ops.quantize_by_tile_backward(data, output_grad, scale, tuple(tile_size), num_bits, offset)
    if doffset.numel() == 0 and doffset.device != data.device:
        doffset = torch.empty(0, device=data.device)
    return [dx, dscale, doffset]


def _cuda_dynamic(data, tile_size, num_bits, symmetric, allow_one_sided, output_dtype):
    return ops.quantize_dynamic_by_tile(data, tuple(tile_size), num_bits, symmetric, allow_one_sided, output_dtype)


def install(fastforward_module: Optional[Any] = None, *, register_linear: bool = True,
            patch_estimators: bool = False) -> dict:
    """Register the B200 kernels with ``fastforward`` (imported if not given).  Idempotent."""
    if _installed:
        return _installed
    if fastforward_module is None:
        import fastforward as fastforward_module  # type: ignore[no-redef]
    ff = fastforward_module
    impl = ff.quantization._quantizer_impl
    table = {
        "quantize_by_tile": (impl.quantize_by_tile_impl, _cuda_quantize),
        "dequantize_by_tile": (impl.dequantize_by_tile_impl, _cuda_dequantize),
        "quantize_by_tile_backward": (impl.quant_dequant_by_tile_grad_impl, _cuda_backward),
        "quantize_dynamic_by_tile": (impl.quantize_dynamic_by_tile_impl, _cuda_dynamic),
    }
    for name, (op_def, kernel) in table.items():
        op_def.register_kernel("cuda")(kernel)       # torch.library.custom_op -> CustomOpDef
        _installed[name] = kernel
    if register_linear:
        _installed["linear"] = _register_linear(ff)
    if patch_estimators:
        from .range_setting import minmax as ours

        ff.range_setting.running_minmax = ours.RunningMinMaxRangeEstimator
        ff.range_setting.minmax.running_minmax = ours.RunningMinMaxRangeEstimator
        _installed["running_minmax"] = ours.RunningMinMaxRangeEstimator
    return _installed


def _register_linear(ff):
    """The reference's QuantizedTensor / params classes differ from ours only by identity, so the
    predicate and kernel are rebuilt against the reference's types."""
    from . import _cabi as C

    QT = ff.QuantizedTensor
    gran = ff.quantization.granularity

    def accepts(input=None, weight=None, bias=None, output_quantizer=None, strict_quantization=None) -> bool:
        if not (isinstance(input, QT) and isinstance(weight, QT)) or isinstance(bias, QT):
            return False
        px, pw = input.quant_args(), weight.quant_args()
        ok = (input.is_cuda and input.raw_data.dtype == torch.int8 and weight.raw_data.dtype == torch.int8
              and weight.dim() == 2 and gran.is_per_tensor(px.granularity) and gran.is_per_channel(pw.granularity)
              and tuple(pw.granularity.channel_dims) == (0,) and weight.shape[1] % 16 == 0
              and getattr(px, "num_bits", 99) <= 8 and getattr(pw, "num_bits", 99) <= 8)
        if not ok:
            return False
        tensors = [px.scale, pw.scale] + [o for o in (px.offset, pw.offset) if o is not None]
        return all(isinstance(t, torch.Tensor) and t.dtype == torch.float32 for t in tensors)

    def kernel(input=None, weight=None, bias=None, output_quantizer=None, strict_quantization=None):
        px, pw = input.quant_args(), weight.quant_args()
        qx = input.raw_data
        k = qx.shape[-1]
        qx2, qw = qx.reshape(-1, k).contiguous(), weight.raw_data.contiguous()
        m, n = qx2.shape[0], qw.shape[0]
        out_dtype = px.dequantize_dtype or torch.float32
        y = torch.empty((m, n), dtype=out_dtype, device=qx.device)
        stream = C.current_stream(qx.device)
        rs_w = torch.empty(n, dtype=torch.int32, device=qx.device)
        C.check(C.lib.ffq_rowsum_i8(qw.data_ptr(), rs_w.data_ptr(), n, k, stream))
        rs_x = None
        if pw.offset is not None:
            rs_x = torch.empty(m, dtype=torch.int32, device=qx.device)
            C.check(C.lib.ffq_rowsum_i8(qx2.data_ptr(), rs_x.data_ptr(), m, k, stream))
        sx, sw = px.scale.detach().reshape(-1), pw.scale.detach().reshape(-1).contiguous()
        ox = None if px.offset is None else px.offset.detach().reshape(-1)
        ow = None if pw.offset is None else pw.offset.detach().reshape(-1).contiguous()
        b = None if bias is None else bias.detach().contiguous()
        ws = torch.empty(4 * n, dtype=torch.float32, device=qx.device)
        C.check(C.lib.ffq_qlinear_w8a8(
            qx2.data_ptr(), qw.data_ptr(), y.data_ptr(), C.dtype_tag(out_dtype), m, n, k, sx.data_ptr(), C.ptr(ox),
            sw.data_ptr(), C.ptr(ow), rs_w.data_ptr(), C.ptr(rs_x), C.ptr(b),
            C.dtype_tag(b.dtype if b is not None else None), ws.data_ptr(), ws.numel() * 4, stream))
        y = y.reshape(*qx.shape[:-1], n)
        return output_quantizer(y) if output_quantizer is not None else y

    return ff.dispatcher.register("linear", ff.dispatcher.Predicate(accepts), kernel)
