"""``disable_quantization`` / ``enable_quantization`` (reference: overrides.py:23-59)."""

from __future__ import annotations

import contextlib
from typing import Any, Generator, List

import torch

from . import flags
from .forward_override import OverrideHandle
from .nn.quantized_module import named_quantizers


class DisableQuantizationOverride:
    def __init__(self) -> None:
        self._enabled = False

    @property
    def quantization_enabled(self) -> bool:
        return self._enabled

    @contextlib.contextmanager
    def enable_quantization(self, enabled: bool = True):
        previous, self._enabled = self._enabled, enabled
        try:
            yield
        finally:
            self._enabled = previous

    def __call__(self, _context: Any, callback, args, kwargs):
        if self._enabled:
            return callback(*args, **kwargs)
        return args[0] if args else next(iter(kwargs.values()))


@contextlib.contextmanager
def disable_quantization(model: torch.nn.Module) -> Generator[None, None, None]:
    handles: List[OverrideHandle] = [q.register_override(DisableQuantizationOverride())
                                     for _, q in named_quantizers(model)]
    try:
        with flags.strict_quantization(False):
            yield
    finally:
        for handle in handles:
            handle.remove()


@contextlib.contextmanager
def enable_quantization(model: torch.nn.Module) -> Generator[None, None, None]:
    with contextlib.ExitStack() as stack:
        for _, quantizer in named_quantizers(model):
            for ov in quantizer.overrides:
                if isinstance(ov, DisableQuantizationOverride):
                    stack.enter_context(ov.enable_quantization())
        yield
