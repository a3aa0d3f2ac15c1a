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
    # the reference computes these parameters with aten's CUDA kernels when its tensors live on a GPU: same flavour
    return ops.quantize_dynamic_by_tile(data, tuple(tile_size), num_bits, symmetric, allow_one_sided, output_dtype,
                                        reciprocal_scalar_division=True)


def install(fastforward_module: Optional[Any] = None, *, register_linear: bool = True,
            patch_estimators: bool = False) -> dict:
    """Register the B200 kernels with ``fastforward`` (imported if not given).  Idempotent."""
    if fastforward_module is None:
        import fastforward as fastforward_module  # type: ignore[no-redef]
    ff = fastforward_module
    if "quantize_by_tile" not in _installed:
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
    if register_linear and "linear" not in _installed:
        _installed["linear"] = _register_linear(ff)
    if patch_estimators:
        install_estimators(ff)
    return _installed


def install_estimators(ff) -> None:
    """Seam 3: ``ff.range_setting.running_minmax`` becomes the sync-free estimator (same name, same arguments).  It
    recognises the reference's own ``LinearQuantizer`` and override chain, so the fused calibration step (one kernel
    per quantizer per forward, no host sync) runs under the unmodified ``ff.estimate_ranges``."""
    from .range_setting import minmax as ours

    if "running_minmax" in _installed:
        return
    _installed["running_minmax_original"] = (ff.range_setting.running_minmax, ff.range_setting.minmax.running_minmax)
    ff.range_setting.running_minmax = ours.RunningMinMaxRangeEstimator
    ff.range_setting.minmax.running_minmax = ours.RunningMinMaxRangeEstimator
    _installed["running_minmax"] = ours.RunningMinMaxRangeEstimator


def uninstall_estimators(ff) -> None:
    orig = _installed.pop("running_minmax_original", None)
    _installed.pop("running_minmax", None)
    if orig is not None:
        ff.range_setting.running_minmax, ff.range_setting.minmax.running_minmax = orig


def _register_linear(ff):
    """The reference's QuantizedTensor / params classes differ from ours only by identity, so the predicates and
    kernels of nn/qlinear.py are rebuilt against the reference's types and registered with ITS dispatcher:
    W8A8 and W4A16 ``linear``, int8 ``matmul`` / ``mm`` / ``bmm`` (dispatcher.py:233-265)."""
    from .nn import qlinear

    kernels = qlinear.build(qlinear.reference_host(ff))
    return qlinear.register_all(kernels, ff.dispatcher.register, ff.dispatcher.Predicate)
