"""Predicate-routed operator registry (reference: dispatcher.py:72-283).

Kernels are kept per operator name in one list ordered by priority; within a priority the most
recent registration is consulted first.  ``dispatch`` calls each predicate with exactly the
arguments of the operator call and returns the first kernel whose predicate accepts them.
This is seam 2 of the drop-in boundary: the W8A8 tensor-core linear is installed with
``register("linear", predicate, kernel)`` (fastforward_b200/nn/qlinear.py)."""

from __future__ import annotations

import enum
import inspect
from typing import Any, Callable, Dict, List, Optional


class _PredicateBase:
    def __call__(self, *args: Any, **kwargs: Any) -> bool:  # pragma: no cover - abstract
        raise NotImplementedError

    def __and__(self, other: "_PredicateBase") -> "_PredicateBase":
        return _Combined(all, (self, other))

    def __or__(self, other: "_PredicateBase") -> "_PredicateBase":
        return _Combined(any, (self, other))

    def __invert__(self) -> "_PredicateBase":
        return _Negated(self)


class _Combined(_PredicateBase):
    def __init__(self, how: Callable[[Any], bool], parts: tuple) -> None:
        self._how, self._parts = how, parts

    def __call__(self, *args: Any, **kwargs: Any) -> bool:
        return self._how(p(*args, **kwargs) for p in self._parts)


class _Negated(_PredicateBase):
    def __init__(self, inner: _PredicateBase) -> None:
        self._inner = inner

    def __call__(self, *args: Any, **kwargs: Any) -> bool:
        return not self._inner(*args, **kwargs)


class Predicate(_PredicateBase):
    """Wraps a boolean function so that predicates compose with ``&``, ``|`` and ``~``."""

    def __init__(self, fn: Callable[..., bool]) -> None:
        self._fn = fn

    def __call__(self, *args: Any, **kwargs: Any) -> bool:
        return bool(self._fn(*args, **kwargs))

    def __repr__(self) -> str:
        try:
            return f"{self._fn.__name__}: {inspect.signature(self._fn)}"
        except (TypeError, ValueError):
            return repr(self._fn)


class DispatcherPriority(enum.IntEnum):
    DEFAULT = 0
    FALLBACK = 1
    NOT_IMPLEMENTED_FALLBACK = 2


class DispatcherItem:
    __slots__ = ("predicate", "fn", "priority")

    def __init__(self, predicate: _PredicateBase, fn: Callable[..., Any], priority: DispatcherPriority) -> None:
        self.predicate, self.fn, self.priority = predicate, fn, priority


_DISPATCHER: Dict[str, List[DispatcherItem]] = {}


class DispatcherRegistrationHook:
    """Returned by ``register``; as a ``with`` target it removes the registration on exit."""

    def __init__(self, op_name: str, item: DispatcherItem) -> None:
        self._op_name, self._item = op_name, item

    def __enter__(self) -> None:
        return None

    def __exit__(self, *exc: Any) -> None:
        self.remove()

    def remove(self) -> None:
        items = _DISPATCHER.get(self._op_name, [])
        if self._item in items:
            items.remove(self._item)


def _always(*_a: Any, **_k: Any) -> bool:
    return True


def register(
    op_name: str,
    predicate: Optional[_PredicateBase] = None,
    kernel: Optional[Callable[..., Any]] = None,
    priority: DispatcherPriority = DispatcherPriority.DEFAULT,
):
    """``register(op, pred, kernel)`` -> hook; ``@register(op, pred)`` -> decorator."""
    if kernel is None:
        def decorator(fn: Callable[..., Any]) -> Callable[..., Any]:
            register(op_name, predicate, fn, priority)
            return fn

        return decorator
    item = DispatcherItem(predicate or Predicate(_always), kernel, priority)
    items = _DISPATCHER.setdefault(op_name, [])
    pos = 0
    while pos < len(items) and items[pos].priority < priority:
        pos += 1          # first slot of this priority class: newest first
    items.insert(pos, item)
    return DispatcherRegistrationHook(op_name, item)


def dispatch(op_name: str, *args: Any, **kwargs: Any) -> Optional[Callable[..., Any]]:
    for item in _DISPATCHER.get(op_name, ()):
        if item.predicate(*args, **kwargs):
            return item.fn
    return None
