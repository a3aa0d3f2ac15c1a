"""``QuantizedModule`` base class, the module conversion map and ``quantize_model``
(reference: nn/quantized_module.py:38-594)."""

from __future__ import annotations

import warnings
from typing import Any, Dict, Iterator, List, Optional, Tuple

import torch

from ..exceptions import QuantizationError
from .quantizer import Quantizer, QuantizerMetadata, QuantizerStub

ModuleConversionDict = Dict[type, Any]
SKIP_QUANTIZATION = object()   # sentinel usable as a conversion target: leave the module as is

_QUANTIZED_MODULE_MAP: Dict[type, List[type]] = {}


def named_quantizers(module: torch.nn.Module, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True,
                     skip_stubs: bool = True) -> Iterator[Tuple[str, Quantizer]]:
    pairs = module.named_modules(prefix="", remove_duplicate=remove_duplicate) if recurse else module.named_children()
    for name, child in pairs:
        if isinstance(child, Quantizer) and not (skip_stubs and isinstance(child, QuantizerStub)):
            yield (f"{prefix}.{name}" if prefix else name), child


def quantizer_state_dict(module: torch.nn.Module) -> Dict[str, Any]:
    state: Dict[str, Any] = {}
    for name, quantizer in named_quantizers(module):
        quantizer.state_dict(destination=state, prefix=f"{name}." if name else "")
    return state


def _record(cls: type) -> None:
    """Associate a QuantizedModule subclass with the single plain nn.Module class it extends."""
    plain = [b for b in cls.__bases__ if issubclass(b, torch.nn.Module) and not issubclass(b, QuantizedModule)]
    if not plain:
        plain = [b for b in cls.__mro__[1:]
                 if issubclass(b, torch.nn.Module) and b is not torch.nn.Module and not issubclass(b, QuantizedModule)]
    if len(plain) == 1:
        _QUANTIZED_MODULE_MAP.setdefault(plain[0], []).append(cls)


class _InitQuantization(type):
    def __call__(cls, *args: Any, **kwargs: Any):
        instance = super().__call__(*args, **kwargs)
        instance.__init_quantization__()
        return instance


class QuantizedModule(torch.nn.Module, metaclass=_InitQuantization):
    """Extension-style base: all quantization set-up lives in ``__init_quantization__`` so that a
    plain module can be converted in place by swapping ``__class__`` and calling it."""

    def __init_quantization__(self) -> None:
        object.__setattr__(self, "_quantizer_metadata", {})

    def __init_subclass__(cls, include_in_module_map: bool = True) -> None:
        if include_in_module_map:
            _record(cls)

    def quantize_children(self, extra_conversion: Optional[ModuleConversionDict] = None,
                          skip_quantized_modules: bool = False, *, ignore_global_module_map: bool = False) -> None:
        for _, child in self.named_children():
            if not isinstance(child, Quantizer):
                quantize_model(child, extra_conversion=extra_conversion, skip_quantized_modules=skip_quantized_modules,
                               ignore_global_module_map=ignore_global_module_map)

    def register_quantizer(self, name: str, quantizer: Optional[Quantizer], *, _register_module: bool = True) -> None:
        if quantizer is not None and not isinstance(quantizer, Quantizer):
            raise TypeError(f"{quantizer} is not a Quantizer subclass")
        slots = self.__dict__.get("_quantizer_metadata")
        if slots is None:
            raise AttributeError(f"Cannot assign quantizer before {type(self).__name__}.__init_quantization__() call")
        if _register_module:
            self.register_module(name, quantizer)
        if quantizer is None:
            return
        # metadata belongs to the module *slot* and is re-attached to whatever is assigned there
        if name not in slots:
            slots[name] = quantizer.quant_metadata if quantizer.quant_metadata is not None else QuantizerMetadata()
        elif quantizer.quant_metadata is not None and not slots[name].is_extension(quantizer.quant_metadata):
            warnings.warn(
                f"Quantizer metadata for {name} is not a consistent extension with stored quantization metadata "
                f"for {name}. The quantizer metadata is updated to match the module.", RuntimeWarning)
        quantizer.quant_metadata = slots[name]

    def __setattr__(self, name: str, value: Any) -> None:
        super().__setattr__(name, value)
        if isinstance(value, Quantizer):
            self.register_quantizer(name, value, _register_module=False)

    def named_quantizers(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True,
                         skip_stubs: bool = True):
        yield from named_quantizers(self, prefix, recurse, remove_duplicate, skip_stubs)

    def quantizers(self, recurse: bool = True, skip_stubs: bool = True):
        for _, q in self.named_quantizers(recurse=recurse, skip_stubs=skip_stubs):
            yield q


def quantized_module_map() -> ModuleConversionDict:
    """Plain module class -> quantized class; the last-defined quantized class wins."""
    result: ModuleConversionDict = {}
    for plain, candidates in _QUANTIZED_MODULE_MAP.items():
        if len(candidates) > 1:
            warnings.warn(f"Multiple quantized implementations for {plain.__name__}; using {candidates[-1].__name__}")
        result[plain] = candidates[-1]
    return result


def _conversion_map(extra: Optional[ModuleConversionDict], ignore_global: bool) -> ModuleConversionDict:
    mapping = {} if ignore_global else quantized_module_map()
    mapping.update(extra or {})
    return mapping


def _unmapped(model: torch.nn.Module, mapping: ModuleConversionDict, skip_quantized: bool, recursive: bool) -> List[type]:
    missing: List[type] = []
    modules = model.modules() if recursive else [model]
    for m in modules:
        if isinstance(m, Quantizer):
            continue
        if isinstance(m, QuantizedModule):
            if skip_quantized:
                continue
            # an already quantized module is fine; it will be skipped or re-initialised
            continue
        if type(m) not in mapping and type(m) not in missing:
            missing.append(type(m))
    return missing


def quantize_model(model: torch.nn.Module, recursive: bool = True, extra_conversion: Optional[ModuleConversionDict] = None,
                   skip_quantized_modules: bool = False, *, ignore_global_module_map: bool = False) -> torch.nn.Module:
    """Convert ``model`` IN PLACE to its quantized counterpart(s): ``module.__class__`` is swapped to
    the registered quantized class and ``__init_quantization__`` is run (quantized_module.py:491-564)."""
    mapping = _conversion_map(extra_conversion, ignore_global_module_map)
    missing = _unmapped(model, mapping, skip_quantized_modules, recursive)
    if missing:
        names = ", ".join(sorted(f"{t.__module__}.{t.__qualname__}" for t in missing))
        raise QuantizationError(
            f"Cannot quantize model because no quantized version of the following modules is known: {names}. "
            "Pass a mapping through `extra_conversion` or implement a QuantizedModule subclass.")
    _convert(model, mapping, recursive, extra_conversion, skip_quantized_modules, ignore_global_module_map)
    return model


def _convert(module, mapping, recursive, extra, skip_quantized, ignore_global) -> None:
    if isinstance(module, QuantizedModule):
        if not skip_quantized:
            pass  # already quantized: keep its quantizers, but still visit children below
        if recursive:
            for child in module.children():
                if not isinstance(child, Quantizer):
                    _convert(child, mapping, recursive, extra, skip_quantized, ignore_global)
        return
    target = mapping[type(module)]
    if target is SKIP_QUANTIZATION:
        return
    module.__class__ = target
    module.__init_quantization__()
    if recursive:
        for child in module.children():
            if not isinstance(child, Quantizer):
                _convert(child, mapping, recursive, extra, skip_quantized, ignore_global)


def surrogate_quantized_modules(model: torch.nn.Module, extra_conversion: Optional[ModuleConversionDict] = None,
                                *, ignore_global_module_map: bool = False) -> ModuleConversionDict:
    """Pass-through ``Quantized<Name>Surrogate`` classes for every module type without a mapping
    (quantized_module.py:422-488): they convert children but add no quantizers themselves."""
    mapping = _conversion_map(extra_conversion, ignore_global_module_map)
    result: ModuleConversionDict = {}
    for m in model.modules():
        t = type(m)
        if isinstance(m, (Quantizer, QuantizedModule)) or t in mapping or t in result:
            continue
        result[t] = type(f"Quantized{t.__name__}Surrogate", (QuantizedModule, t), {}, include_in_module_map=False)
    return result
