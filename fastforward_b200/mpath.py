"""Module-path queries: a small implementation of the query language the reference uses to
select modules (reference: mpath/__init__.py:53-107, mpath/_parser.py:311-407).

Supported grammar (what ``find_quantizers`` needs): fragments separated by ``/``; ``name`` or digits =
exact child name; ``*`` = exactly one level; ``**`` = zero or more levels; ``[cls:dotted.Name]`` /
``[class:...]`` = isinstance test (name resolved in the caller's globals/locals, then by import);
``[re:pattern]`` / ``[regex:...]`` = full match on the child name; ``[quantizer:tag1,tag2]`` /
``[qtag:...]`` = a Quantizer carrying all tags; ``~fragment`` = negation; ``{a, b}`` = alternatives.
The root never matches; every module is reported once.  Host-only string/tree matching."""

from __future__ import annotations

import dataclasses
import importlib
import re
import sys
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional

import torch


@dataclasses.dataclass
class FilterResult:
    full_name: str
    module: torch.nn.Module
    parent: Optional[torch.nn.Module]
    parent_attribute: str

    def update_module(self, new_module: torch.nn.Module, safe: bool = True) -> "FilterResult":
        if self.parent is None:
            raise ValueError("Cannot replace the root module")
        if safe and getattr(self.parent, self.parent_attribute, None) is not self.module:
            raise RuntimeError(
                f"'{self.full_name}' was replaced after this result was created; pass safe=False to force")
        setattr(self.parent, self.parent_attribute, new_module)
        return FilterResult(self.full_name, new_module, self.parent, self.parent_attribute)


Matcher = Callable[[str, torch.nn.Module], bool]


def _split_top(text: str, sep: str) -> List[str]:
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "[{":
            depth += 1
        elif ch in "]}":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def _resolve_class(name: str, frame_vars: Dict[str, Any]) -> Any:
    head, *rest = name.split(".")
    obj = frame_vars.get(head)
    if obj is None:
        try:
            obj = importlib.import_module(head)
        except ImportError:
            raise ValueError(f"Cannot resolve class '{name}' in mpath query") from None
    for attr in rest:
        if hasattr(obj, attr):
            obj = getattr(obj, attr)
        else:
            obj = importlib.import_module(f"{obj.__name__}.{attr}")
    return obj


def _parse_fragment(text: str, frame_vars: Dict[str, Any]) -> Matcher:
    text = text.strip()
    if text.startswith("~"):
        inner = _parse_fragment(text[1:], frame_vars)
        return lambda n, m: not inner(n, m)
    if text.startswith("{") and text.endswith("}"):
        options = [_parse_fragment(t, frame_vars) for t in _split_top(text[1:-1], ",")]
        return lambda n, m: any(o(n, m) for o in options)
    if text == "*":
        return lambda n, m: True
    if text.startswith("[") and text.endswith("]"):
        kind, _, spec = text[1:-1].partition(":")
        kind, spec = kind.strip(), spec.strip()
        if kind in ("cls", "class"):
            cls = _resolve_class(spec, frame_vars)
            return lambda n, m: isinstance(m, cls)
        if kind in ("re", "regex"):
            pattern = re.compile(spec)
            return lambda n, m: pattern.fullmatch(n) is not None
        if kind in ("quantizer", "qtag"):
            from .nn.quantizer import Quantizer, Tag

            tags = [Tag(t.strip()) for t in spec.split(",") if t.strip()]

            def has_tags(n: str, m: torch.nn.Module) -> bool:
                meta = getattr(m, "quant_metadata", None)
                return isinstance(m, Quantizer) and meta is not None and all(t in meta for t in tags)

            return has_tags
        raise ValueError(f"Unknown mpath extension '{kind}'")
    return lambda n, m: n == text


def _search(root: torch.nn.Module, fragments: List[Any]) -> List[FilterResult]:
    results: List[FilterResult] = []
    seen = set()

    def visit(module: torch.nn.Module, name: str, idx: int) -> None:
        if idx == len(fragments):
            return
        frag = fragments[idx]
        if frag == "**":
            visit(module, name, idx + 1)                       # zero levels
            for child_name, child in module.named_children():  # one more level, stay on '**'
                visit(child, f"{name}.{child_name}" if name else child_name, idx)
            return
        for child_name, child in module.named_children():
            if not frag(child_name, child):
                continue
            full = f"{name}.{child_name}" if name else child_name
            if idx + 1 == len(fragments):
                if id(child) not in seen:
                    seen.add(id(child))
                    results.append(FilterResult(full, child, module, child_name))
            else:
                visit(child, full, idx + 1)

    if fragments and fragments[-1] == "**":
        fragments = fragments + [lambda n, m: True]
    visit(root, "", 0)
    return results


def search(query: str, root: torch.nn.Module, *, _frame_depth: int = 1, aliases: Optional[Dict[str, str]] = None):
    frame = sys._getframe(_frame_depth)
    frame_vars = {**frame.f_globals, **frame.f_locals}
    query = query.strip()
    if query.startswith("/"):
        query = query[1:]
    for alias, value in (aliases or {}).items():
        query = query.replace(f"&{alias}", value)
    fragments = []
    for part in _split_top(query, "/"):
        part = part.strip()
        if not part:
            continue
        fragments.append("**" if part == "**" else _parse_fragment(part, frame_vars))
    return MPathCollection(root, _search(root, fragments))


class MPathCollection:
    def __init__(self, root: torch.nn.Module, results: Optional[Iterable[FilterResult]] = None) -> None:
        self._root = root
        self._results: List[FilterResult] = list(results or [])

    def __len__(self) -> int:
        return len(self._results)

    def __iter__(self) -> Iterator[FilterResult]:
        return iter(self._results)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return type(self)(self._root, self._results[idx])
        return self._results[idx]

    def append(self, item: FilterResult) -> None:
        self._results.append(item)

    def modules(self) -> Iterator[torch.nn.Module]:
        return (r.module for r in self._results)

    def named_modules(self):
        return ((r.full_name, r.module) for r in self._results)

    def parents(self):
        return (r.parent for r in self._results)

    def apply(self, fn: Callable[[torch.nn.Module], Any]) -> "MPathCollection":
        for r in self._results:
            fn(r.module)
        return self

    def map(self, fn: Callable[[str, torch.nn.Module], torch.nn.Module]) -> "MPathCollection":
        self._results = [r.update_module(fn(r.full_name, r.module)) for r in self._results]
        return self

    def _combine(self, other: "MPathCollection", keep: Callable[[bool, bool], bool]) -> "MPathCollection":
        mine = {id(r.module): r for r in self._results}
        theirs = {id(r.module): r for r in other._results}
        out = [r for k, r in {**theirs, **mine}.items() if keep(k in mine, k in theirs)]
        return type(self)(self._root, out)

    def __or__(self, other):
        return self._combine(other, lambda a, b: a or b)

    def __and__(self, other):
        return self._combine(other, lambda a, b: a and b)

    def __sub__(self, other):
        return self._combine(other, lambda a, b: a and not b)

    def __xor__(self, other):
        return self._combine(other, lambda a, b: a != b)
