"""Freeze quantized parameter values into a model during one forward pass
(reference: quantization/freeze.py:14-125).

``with freeze_parameters(model): model(batch)`` registers an override on every quantizer; when a quantizer is handed an
``nn.Parameter`` the parameter is replaced IN PLACE by ``quantizer(param).dequantize()`` and (by default) the quantizer
is replaced by a ``QuantizerStub`` carrying its metadata.  The reference does this with a quantize chain, a dequantize
chain and a ``copy_`` (>= 5 passes over the weight plus three full-size temporaries); here a calibrated
``LinearQuantizer`` with nothing else overriding it takes the fused fake-quantize kernel writing straight into the
parameter's storage -- one read and one write per element, bit-identical (``quantization/fuse.py: _fused_inplace``).
Every other case (other quantizer types, further overrides such as ``disable_quantization``, non-Parameter inputs)
follows the reference's sequence literally."""

from __future__ import annotations

import contextlib
from typing import Any, Callable, Iterator, List, Sequence, Union

import torch

from .. import flags
from ..nn.linear_quantizer import LinearQuantizer
from ..nn.quantized_module import named_quantizers
from ..nn.quantizer import QuantizerStub

_stats = {"fused_in_place": 0, "two_step": 0}


class _FreezeParametersOverride:
    def __init__(self, module: torch.nn.Module, remove_quantizer: bool) -> None:
        self._module = module
        self._remove_quantizer = remove_quantizer

    def __call__(self, quantizer, callback: Callable[..., Any], args: tuple, kwargs: dict) -> torch.Tensor:
        from ..range_setting.minmax import _next_is_own_quantize
        from .fuse import _fused_inplace

        input_data = args[0] if args else next(iter(kwargs.values()))
        done = False
        if isinstance(input_data, torch.nn.Parameter) and type(quantizer) is LinearQuantizer \
                and _next_is_own_quantize(callback, quantizer) and not flags.get_export_mode():
            done = _fused_inplace(input_data, quantizer)       # elementwise, so writing in place is safe
        if done:
            _stats["fused_in_place"] += 1
        else:
            quantized = callback(input_data).dequantize()
            if quantized is input_data:
                # a no-op quantizer (e.g. disabled): nothing to freeze, nothing to remove (freeze.py:44-49)
                return input_data
            if isinstance(input_data, torch.nn.Parameter):
                with torch.no_grad():
                    input_data.copy_(quantized)
            _stats["two_step"] += 1
        if self._remove_quantizer:
            for name, other in named_quantizers(self._module, recurse=False, skip_stubs=False):
                if other is quantizer:
                    setattr(self._module, name, QuantizerStub(_metadata=quantizer.quant_metadata))
                    break
        return input_data


@contextlib.contextmanager
def freeze_parameters(modules: Union[torch.nn.Module, Sequence[torch.nn.Module]],
                      remove_quantizers: bool = True) -> Iterator[None]:
    """See the module docstring; the caller runs a single forward pass inside the block.  Strict quantization is
    disabled inside it because the overrides return plain tensors; the overrides are removed on exit."""
    hooks: List[Any] = []
    modules = [modules] if isinstance(modules, torch.nn.Module) else modules
    for module in modules:
        for submodule in module.modules():
            for _, quantizer in named_quantizers(submodule, recurse=False):
                hooks.append(quantizer.register_override(_FreezeParametersOverride(submodule, remove_quantizers)))
    try:
        with flags.strict_quantization(False):
            yield
    finally:
        for hook in hooks:
            hook.remove()


def stats() -> dict:
    return dict(_stats)
