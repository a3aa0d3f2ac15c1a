"""YAML form of quantizers and granularities: the on-disk configuration of a quantized model
(reference: serialization.py:62-220, used by quantization/save_load.py for ``config.yaml``).

The FORMAT is the reference's, so files travel in both directions between this package and an installed
``fastforward``: an object is a mapping tagged ``!ff.obj`` with

    name      fully qualified class name, always spelled with the reference's root package (``fastforward.``)
    initargs  ``(args, kwargs)`` of the most derived ``__init__``
    state     ``__getstate__()`` (for quantizers: the non-tensor attributes)

and a ``torch.dtype`` is ``!ff.obj {name: torch.int8}``.  The MECHANISM is this package's own: constructor arguments
are captured by ``remember_init_args`` (a class decorator that also covers later subclasses), and the YAML hooks live
on a private ``Dumper`` / ``Loader`` pair instead of PyYAML's global ones, so that importing this package next to the
reference (``plugin.install()``) does not fight over the ``!ff.obj`` tag.  Host-only: no device work here."""

from __future__ import annotations

import copy
import functools
import importlib
from typing import Any

import torch
import yaml

_OWN_ROOT = __name__.partition(".")[0]
_REF_ROOT = "fastforward"
_ARGS_ATTR = "_ffq_init_args"
TAG = "!ff.obj"


# ---------------------------------------------------------------------------------------------------------------
# constructor-argument capture
# ---------------------------------------------------------------------------------------------------------------
def _capturing(init):
    @functools.wraps(init)
    def wrapped(self, *args: Any, **kwargs: Any) -> None:
        captured = copy.deepcopy((tuple(args), dict(kwargs)))
        init(self, *args, **kwargs)
        # the most derived __init__ returns last, so its arguments are the ones that stay
        object.__setattr__(self, _ARGS_ATTR, captured)

    wrapped._ffq_capturing = True
    return wrapped


def _wrap_own_init(cls: type) -> None:
    init = cls.__dict__.get("__init__")
    if init is not None and not getattr(init, "_ffq_capturing", False):
        cls.__init__ = _capturing(init)


def remember_init_args(cls: type) -> type:
    """Class decorator: instances of ``cls`` and of every subclass defined later remember the arguments their most
    derived ``__init__`` was called with -- what ``dump`` writes as ``initargs`` and ``load`` constructs from."""
    _wrap_own_init(cls)
    inherited = cls.__dict__.get("__init_subclass__")

    def __init_subclass__(sub, **kwargs: Any) -> None:
        if inherited is not None:
            inherited.__func__(sub, **kwargs)
        else:
            super(cls, sub).__init_subclass__(**kwargs)
        _wrap_own_init(sub)

    cls.__init_subclass__ = classmethod(__init_subclass__)
    _SERIALIZABLE.append(cls)
    _Dumper.add_multi_representer(cls, _represent_object)
    return cls


def init_args(obj: Any):
    """``(args, kwargs)`` the object was constructed with, or None when it was not created through ``__init__``."""
    return getattr(obj, _ARGS_ATTR, None)


# ---------------------------------------------------------------------------------------------------------------
# names: written with the reference's root, resolved in this package first
# ---------------------------------------------------------------------------------------------------------------
def portable_name(cls: type) -> str:
    module = cls.__module__
    if module == _OWN_ROOT or module.startswith(_OWN_ROOT + "."):
        module = _REF_ROOT + module[len(_OWN_ROOT):]
    return f"{module}.{cls.__qualname__}"


def _import_dotted(name: str) -> Any:
    parts = name.split(".")
    for cut in range(len(parts) - 1, 0, -1):
        try:
            obj = importlib.import_module(".".join(parts[:cut]))
        except ImportError:
            continue
        try:
            for attr in parts[cut:]:
                obj = getattr(obj, attr)
        except AttributeError:
            continue
        return obj
    raise ImportError(f"cannot resolve '{name}'")


def resolve_name(name: str) -> Any:
    """The object a ``name`` entry refers to.  ``fastforward.*`` resolves to this package's class of the same
    path (the file may have been written by either package); anything else is imported as spelled."""
    if name == _REF_ROOT or name.startswith(_REF_ROOT + "."):
        try:
            return _import_dotted(_OWN_ROOT + name[len(_REF_ROOT):])
        except ImportError:
            pass
    return _import_dotted(name)


# ---------------------------------------------------------------------------------------------------------------
# YAML hooks on a private Dumper / Loader
# ---------------------------------------------------------------------------------------------------------------
class _Dumper(yaml.Dumper):
    pass


class _Loader(yaml.Loader):
    pass


_SERIALIZABLE: list = []


def _represent_object(dumper: yaml.Dumper, obj: Any) -> yaml.Node:
    args = init_args(obj)
    if args is None:
        if type(obj).__init__ is not object.__init__:
            raise RuntimeError(f"{type(obj).__name__} was not constructed through __init__: its arguments are unknown "
                               "and it cannot be written to YAML")
        args = ((), {})                                  # a class without a constructor of its own (PerTensor)
    node = {"name": portable_name(type(obj)), "initargs": (tuple(args[0]), dict(args[1]))}
    if hasattr(obj, "_yaml_state"):
        node["state"] = obj._yaml_state()
    elif hasattr(obj, "__setstate__"):
        state = obj.__getstate__()
        if isinstance(state, dict):
            node["state"] = state
    return dumper.represent_mapping(TAG, node)


def _represent_dtype(dumper: yaml.Dumper, dtype: torch.dtype) -> yaml.Node:
    return dumper.represent_mapping(TAG, {"name": str(dtype)})


def _construct(loader: yaml.Loader, _suffix: str, node: yaml.Node) -> Any:
    fields = loader.construct_mapping(node, deep=True)
    target = resolve_name(fields.pop("name"))
    if not isinstance(target, type):
        return target                                   # a torch.dtype (or any other named constant)
    new_args, new_kwargs = fields.get("newargs", ((), {}))
    args, kwargs = fields.get("initargs", ((), {}))
    if "initargs" in fields or "newargs" not in fields:
        obj = target(*args, **kwargs)
    else:
        obj = target.__new__(target, *new_args, **new_kwargs)
    state = fields.get("state")
    if state is not None:
        if hasattr(obj, "_yaml_setstate"):
            obj._yaml_setstate(state)
        elif hasattr(obj, "__setstate__"):
            obj.__setstate__(state)
        else:
            obj.__dict__.update(state)
    return obj


_Dumper.add_multi_representer(torch.dtype, _represent_dtype)
_Loader.add_multi_constructor(TAG, _construct)


def dump(data: Any, stream=None, **kwargs: Any):
    """``yaml.dump`` with the ``!ff.obj`` representers."""
    kwargs.setdefault("sort_keys", False)
    return yaml.dump(data, stream, Dumper=_Dumper, **kwargs)


def load(stream) -> Any:
    """``yaml.load`` with the ``!ff.obj`` constructor (full Python loader, as the reference uses: only load files you
    trust)."""
    return yaml.load(stream, Loader=_Loader)
