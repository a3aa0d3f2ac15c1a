"""``find_quantizers`` / ``QuantizerCollection.initialize`` / ``QuantizationConfig``
(reference: quantization/quant_init.py:72-393)."""

from __future__ import annotations

from typing import Any, Callable, Dict, Iterable, List, Optional, Tuple

import torch

from . import mpath
from .exceptions import QuantizationError
from .nn.quantizer import Quantizer, QuantizerStub


def _factory_fn(factory, kwargs: Dict[str, Any]) -> Callable[[str, Quantizer], Quantizer]:
    if isinstance(factory, type):
        if not issubclass(factory, Quantizer):
            raise TypeError(f"{factory} is not a Quantizer subclass")
        return lambda _name, _existing: factory(**kwargs)
    return factory


def _initialize_one(result: mpath.FilterResult, factory, overwrite_policy: str, safe: bool) -> mpath.FilterResult:
    if not isinstance(result.module, Quantizer):
        raise TypeError(f"'{result.full_name}' is not a quantizer.")
    if not isinstance(result.module, QuantizerStub):
        if overwrite_policy == "error":
            raise QuantizationError(
                f"'{result.full_name}' is a quantizer, but is already initialized. If you want to overwrite the "
                'existing quantizer, use overwrite_policy="overwrite" or if you want to skip re-initializing '
                'existing quantizers use overwrite_policy="skip"')
        if overwrite_policy == "skip":
            return result
        if overwrite_policy != "overwrite":
            raise ValueError(
                f"Overwrite would occur, but overwrite_policy={overwrite_policy!r} is illegal. please use one of "
                '"overwrite", "skip", or "error"')
    return result.update_module(factory(result.full_name, result.module), safe=safe)


class QuantizerCollection(mpath.MPathCollection):
    """A collection that only holds Quantizer results."""

    def __init__(self, root: torch.nn.Module, results: Optional[Iterable[mpath.FilterResult]] = None) -> None:
        super().__init__(root, [r for r in (results or []) if isinstance(r.module, Quantizer)])

    def append(self, item: mpath.FilterResult) -> None:
        if not isinstance(item.module, Quantizer):
            raise ValueError(f"Can only insert a FilterResult of a Quantizer module to {type(self).__name__}")
        super().append(item)

    def initialize(self, quantizer_factory, *, overwrite_policy: str = "error", safe: bool = True, **kwargs: Any) -> None:
        factory = _factory_fn(quantizer_factory, kwargs)
        self._results = [_initialize_one(r, factory, overwrite_policy, safe) for r in self._results]


def find_quantizers(root: torch.nn.Module, query: str, *, aliases: Optional[Dict[str, str]] = None) -> QuantizerCollection:
    """``find_quantizers(model, "**/[quantizer:parameter/weight]")`` (quant_init.py:214-236)."""
    found = mpath.search(query, root, _frame_depth=2, aliases=aliases)
    return QuantizerCollection(root, list(found))


class QuantizationConfig:
    """Ordered ``(query, factory)`` rules; for every quantizer the LAST matching rule wins
    (quant_init.py:277-393)."""

    def __init__(self) -> None:
        self._rules: List[Tuple[str, Callable[[str, Quantizer], Quantizer]]] = []

    def add_rule(self, query: str, quantizer_factory, **kwargs: Any) -> "QuantizationConfig":
        self._rules.append((query, _factory_fn(quantizer_factory, kwargs)))
        return self

    def initialize(self, model: torch.nn.Module, *, overwrite_policy: str = "error", safe: bool = True) -> None:
        chosen: Dict[int, Tuple[mpath.FilterResult, Any]] = {}
        for query, factory in self._rules:
            for result in QuantizerCollection(model, list(mpath.search(query, model, _frame_depth=2))):
                chosen[id(result.module)] = (result, factory)
        for result, factory in chosen.values():
            _initialize_one(result, factory, overwrite_policy, safe)
