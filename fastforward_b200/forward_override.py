"""Override stack around ``Quantizer.quantize`` (reference: forward_override.py:16-125).

Overrides are keyed by a globally increasing handle id.  On every forward the stack is rebuilt
and unwound newest-first: the most recently registered override runs outermost and receives
``(quantizer, next_fn, args, kwargs)``; calling ``next_fn`` continues with the next older one and
finally the quantizer's own ``quantize``."""

from __future__ import annotations

import itertools
import weakref
from typing import Any, Callable, Mapping, Optional

_handle_ids = itertools.count()


class OverrideHandle:
    def __init__(self, quantizer: Any) -> None:
        self._quantizer = weakref.ref(quantizer)
        self.handle_id = next(_handle_ids)

    def remove(self) -> Optional[Callable[..., Any]]:
        quantizer = self._quantizer()
        return None if quantizer is None else quantizer.remove_override(self.handle_id)

    def __enter__(self) -> "OverrideHandle":
        return self

    def __exit__(self, *exc: Any) -> None:
        self.remove()


class _Chain:
    def __init__(self, context: Any, fn: Callable[..., Any], overrides: Mapping[int, Callable[..., Any]]) -> None:
        self._context, self._fn = context, fn
        self._pending = [overrides[k] for k in sorted(overrides)]

    def __call__(self, *args: Any, **kwargs: Any) -> Any:
        if not self._pending:
            return self._fn(*args, **kwargs)
        return self._pending.pop()(self._context, self, args, kwargs)


def apply_overrides(context: Any, overridden_fn: Callable[..., Any], override_map: Mapping[int, Callable[..., Any]]):
    return _Chain(context, overridden_fn, override_map) if override_map else overridden_fn
