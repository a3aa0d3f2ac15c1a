"""Quantization parameters / function / context triple (reference: quantization/function.py:24-166).

A ``QuantizationContext`` is what a ``QuantizedTensor`` carries: the function class that knows how
to dequantize the raw codes plus the parameter dataclass (scale, offset, granularity, ...)."""

from __future__ import annotations

import abc
import dataclasses
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
