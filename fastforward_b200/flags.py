"""Process-global behaviour flags with the reference's names and set/get/context semantics
(flags.py:28-106 in the reference): ``set_X(v)`` applies immediately AND returns a context
manager that restores the previous value on exit."""

from __future__ import annotations

import functools
from typing import Any, Callable

_FLAGS = {
    "strict_quantization": True,
    "export_mode": False,
    "compiled_quant_funcs": False,   # meaningless here: the CUDA kernels are already fused
    "sdpa_torch_fallback_allowed": False,
}


class _Restore:
    def __init__(self, name: str, previous: bool) -> None:
        self._name, self._previous = name, previous

    def __enter__(self) -> None:
        return None

    def __exit__(self, *exc: Any) -> None:
        _FLAGS[self._name] = self._previous


def _setter(name: str) -> Callable[[bool], _Restore]:
    def set_flag(value: bool) -> _Restore:
        previous = _FLAGS[name]
        _FLAGS[name] = bool(value)
        return _Restore(name, previous)

    set_flag.__name__ = f"set_{name}"
    return set_flag


def _getter(name: str) -> Callable[[], bool]:
    def get_flag() -> bool:
        return _FLAGS[name]

    get_flag.__name__ = f"get_{name}"
    return get_flag


class _FlagContext:
    """``with strict_quantization(False): ...`` -- also usable as a decorator."""

    _name = ""

    def __init__(self, value: bool) -> None:
        self._value = bool(value)
        self._stack: list[bool] = []

    def __enter__(self) -> None:
        self._stack.append(_FLAGS[self._name])
        _FLAGS[self._name] = self._value

    def __exit__(self, *exc: Any) -> None:
        _FLAGS[self._name] = self._stack.pop()

    def __call__(self, fn: Callable[..., Any]) -> Callable[..., Any]:
        @functools.wraps(fn)
        def wrapper(*a: Any, **k: Any) -> Any:
            with type(self)(self._value):
                return fn(*a, **k)

        return wrapper


def _context(name: str) -> type:
    return type(name, (_FlagContext,), {"_name": name})


set_strict_quantization, get_strict_quantization = _setter("strict_quantization"), _getter("strict_quantization")
set_export_mode, get_export_mode = _setter("export_mode"), _getter("export_mode")
set_compiled_quant_funcs, get_compiled_quant_funcs = _setter("compiled_quant_funcs"), _getter("compiled_quant_funcs")
set_sdpa_torch_fallback_allowed = _setter("sdpa_torch_fallback_allowed")
get_sdpa_torch_fallback_allowed = _getter("sdpa_torch_fallback_allowed")
strict_quantization = _context("strict_quantization")
export_mode = _context("export_mode")
compiled_quant_funcs = _context("compiled_quant_funcs")
