"""Range-estimation plumbing (reference: range_setting/common.py:29-289)."""

from __future__ import annotations

import abc
import contextlib
from typing import Any, Callable, Generator, Iterator, Protocol, Sequence, runtime_checkable

import torch


@runtime_checkable
class RangeSettable(Protocol):
    granularity: Any

    @property
    def quantization_range(self): ...

    @quantization_range.setter
    def quantization_range(self, value) -> None: ...


class RangeEstimator(abc.ABC):
    @abc.abstractmethod
    def prepare(self, module) -> Any: ...

    @abc.abstractmethod
    def cleanup(self, module, metadata) -> None: ...

    @abc.abstractmethod
    def split_module(self, module: torch.nn.Module) -> Iterator[Any]: ...

    def finalize(self, prepared: Sequence[tuple]) -> None:
        """Called once when the ``estimate_ranges`` block ends, before cleanup.  The B200 backend
        uses it for what the reference does inline with host syncs: the deferred +-inf check and
        the multi-GPU range all-reduce."""


class SimpleEstimatorStep(abc.ABC):
    """Override callable ``(quantizer, callback, args, kwargs) -> Tensor``: runs ``estimate_step`` on
    the data and then either forwards to the quantizer or returns the data unquantized."""

    def __init__(self, *args: Any, disable_quantization: bool = False, **kwargs: Any) -> None:
        self._initialized = False
        self._disable_quantization = disable_quantization
        super().__init__(*args, **kwargs)

    def setup_estimator(self, data: torch.Tensor) -> None:
        pass

    @abc.abstractmethod
    def estimate_step(self, quantizer, data: torch.Tensor) -> None: ...

    def forward(self, quantizer, callback: Callable[[torch.Tensor], torch.Tensor], args: tuple, kwargs: dict):
        data = args[0] if args else kwargs.get("data", next(iter(kwargs.values()), None))
        if not self._initialized:
            self.setup_estimator(data)
            self._initialized = True
        self.estimate_step(quantizer, data)
        return data if self._disable_quantization else callback(data)


@contextlib.contextmanager
def estimate_ranges(model_or_layers, estimator, *args: Any, **kwargs: Any) -> Generator[None, None, None]:
    """``with estimate_ranges(model, running_minmax): model(batch)`` (common.py:241-289)."""
    if isinstance(model_or_layers, torch.nn.Module):
        model_or_layers = [model_or_layers]
    if isinstance(estimator, type):
        estimator = estimator(*args, **kwargs)
    elif args or kwargs:
        raise ValueError("`estimator` is already initialized so no `args` or `kwargs` can be given.")
    prepared = []
    for module in model_or_layers:
        for part in estimator.split_module(module):
            prepared.append((part, estimator.prepare(part)))
    failed = True
    cleaned = []

    def cleanup_all() -> None:
        if not cleaned:
            cleaned.append(True)
            for module, metadata in prepared:
                estimator.cleanup(module, metadata)

    try:
        yield
        failed = False
    finally:
        try:
            if not failed and hasattr(estimator, "finalize"):
                if getattr(estimator, "finalize_runs_cleanup", False):
                    # the estimator removes the overrides itself, BEFORE its one host sync: host work that needs nothing
                    # from the device belongs in front of the wait, not behind it
                    estimator.finalize(prepared, cleanup=cleanup_all)
                else:
                    estimator.finalize(prepared)
        finally:
            cleanup_all()
