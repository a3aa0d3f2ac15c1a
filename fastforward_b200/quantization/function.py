"""Quantization parameters / function / context triple (reference: quantization/function.py:24-166).

A ``QuantizationContext`` is what a ``QuantizedTensor`` carries: the function class that knows how
to dequantize the raw codes plus the parameter dataclass (scale, offset, granularity, ...)."""

from __future__ import annotations

import abc
import dataclasses
import inspect
from typing import Any, Callable, Generic, TypeVar

import torch

from .. import flags


def _tensor_apply(fn: Callable[[torch.Tensor], torch.Tensor]) -> Callable[[Any], Any]:
    return lambda v: fn(v) if isinstance(v, torch.Tensor) else v


@dataclasses.dataclass
class QuantizationParameters:
    """Base class of parameter records.  Fields are shared by reference, never deep-copied."""

    def _fields(self) -> dict:
        return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}

    def with_changes(self, **changes: Any):
        return dataclasses.replace(self, **changes)

    def _apply(self, fn: Callable[[Any], Any]):
        return type(self)(**{k: fn(v) for k, v in self._fields().items()})

    def __format__(self, spec: str) -> str:
        return repr(self)


P = TypeVar("P", bound=QuantizationParameters)


class QuantizationFunction(abc.ABC, Generic[P]):
    @classmethod
    @abc.abstractmethod
    def quantize(cls, data: torch.Tensor, params: P):
        """Return ``data`` quantized following ``params`` as a QuantizedTensor."""

    @classmethod
    @abc.abstractmethod
    def dequantize(cls, data: torch.Tensor, params: P) -> torch.Tensor:
        """Return the real-valued tensor represented by raw codes ``data``."""


@dataclasses.dataclass(frozen=True)
class QuantizationContext(Generic[P]):
    quantization_fn: type
    quantization_params: Any

    def with_changes(self, quantization_fn: type | None = None, **changes: Any):
        params = self.quantization_params.with_changes(**changes)
        return QuantizationContext(quantization_fn or self.quantization_fn, params)

    def _apply(self, fn: Callable[[Any], Any]):
        return dataclasses.replace(self, quantization_params=self.quantization_params._apply(fn))

    def clone_parameters(self):
        return self._apply(_tensor_apply(torch.clone))

    def detach_parameters(self):
        return self._apply(_tensor_apply(torch.detach))

    def contiguous_parameters(self):
        new = self._apply(_tensor_apply(torch.Tensor.contiguous))
        old_f, new_f = self.quantization_params._fields(), new.quantization_params._fields()
        return new if any(old_f[k] is not new_f[k] for k in old_f) else self

    def to(self, device):
        return self._apply(_tensor_apply(lambda t: t.to(device=device)))

    def attach(self, data: torch.Tensor):
        """Wrap raw codes ``data`` as a QuantizedTensor of this context (a plain dequantized
        tensor in export mode, function.py:162-165)."""
        if flags.get_export_mode():
            return self.quantization_fn.dequantize(data, self.quantization_params)
        from ..quantized_tensor import QuantizedTensor

        return QuantizedTensor(data, self)


# ---------------------------------------------------------------------------------------------------------------
# a QuantizationFunction from a pair of plain functions (reference: quantization/function.py:209-330)
# ---------------------------------------------------------------------------------------------------------------
_NO_DEFAULT = inspect.Parameter.empty


def _keyword_parameters(fn: Callable[..., torch.Tensor]) -> dict:
    """name -> (annotation, default) of every parameter after the data tensor; all must be passable by keyword."""
    params = list(inspect.signature(fn).parameters.values())[1:]
    out = {}
    for p in params:
        if p.kind not in (p.KEYWORD_ONLY, p.POSITIONAL_OR_KEYWORD):
            raise TypeError(f"All parameters must be keyword only or positional or keyword parameters {p.name} is "
                            f"{p.kind.description}")
        out[p.name] = (Any if p.annotation is _NO_DEFAULT else p.annotation, p.default)
    return out


def create_quantization_function(cls_name: str, quantize: Callable[..., torch.Tensor],
                                 dequantize: Callable[..., torch.Tensor]):
    """``(ParamsType, FunctionType, helper)`` for a custom quantizer given as two plain functions
    ``quantize(data, **p) -> Tensor`` and ``dequantize(data, **p) -> Tensor``.  ``ParamsType`` is a dataclass holding
    the union of their keyword parameters (a parameter both share must agree in annotation and default),
    ``FunctionType`` a ``QuantizationFunction`` whose ``quantize`` wraps the codes in a ``QuantizedTensor`` carrying
    those parameters, and ``helper(data, **p)`` builds the parameters and quantizes in one call."""
    q_params, d_params = _keyword_parameters(quantize), _keyword_parameters(dequantize)
    for name in q_params.keys() & d_params.keys():
        if q_params[name][0] != d_params[name][0]:
            raise TypeError(f"The type annotation for '{name}' must be the same in both the quantize and dequantize function")
        if q_params[name][1] != d_params[name][1]:
            raise ValueError(f"The default value for '{name}' must be the same in both the quantize and dequantize function")
    merged = {**q_params, **d_params}
    fields = [(name, annotation, dataclasses.field() if default is _NO_DEFAULT else dataclasses.field(default=default))
              for name, (annotation, default) in merged.items()]
    params_type = dataclasses.make_dataclass(
        f"{cls_name}Params", fields, bases=(QuantizationParameters,),
        namespace={"quantize_params": lambda self: {n: getattr(self, n) for n in q_params},
                   "dequantize_params": lambda self: {n: getattr(self, n) for n in d_params}})

    def _quantize(cls, data: torch.Tensor, params):
        from ..quantized_tensor import QuantizedTensor

        return QuantizedTensor(quantize(data, **params.quantize_params()), QuantizationContext(cls, params))

    def _dequantize(cls, data: torch.Tensor, params) -> torch.Tensor:
        return dequantize(data, **params.dequantize_params())

    function_type = type(cls_name, (QuantizationFunction,), {"quantize": classmethod(_quantize),
                                                               "dequantize": classmethod(_dequantize)})

    def helper(data: torch.Tensor, **kwargs: Any):
        return function_type.quantize(data, params_type(**kwargs))

    return params_type, function_type, helper
