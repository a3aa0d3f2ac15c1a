"""Quantizer base class, stubs, tags and metadata (reference: nn/quantizer.py:22-535)."""

from __future__ import annotations

import collections
import copy
from typing import Any, Callable, Dict, Iterator, Optional

import torch

from .. import forward_override as override
from ..serialization import remember_init_args


class Tag:
    """Interned, '/'-hierarchical symbol: ``Tag("parameter/weight")``."""

    _tags: Dict[str, "Tag"] = {}

    def __new__(cls, symbol):
        if isinstance(symbol, Tag):
            return symbol
        tag = cls._tags.get(symbol)
        if tag is None:
            tag = super().__new__(cls)
            tag._symbol = symbol
            cls._tags[symbol] = tag
        return tag

    def __copy__(self):
        return self

    def __deepcopy__(self, memo):
        return self

    def __str__(self) -> str:
        return f"#{self._symbol}"

    def __repr__(self) -> str:
        return self._symbol

    def hierarchy(self) -> Iterator["Tag"]:
        parts = self._symbol.split("/")
        for n in range(len(parts), 0, -1):
            yield Tag("/".join(parts[:n]))

    def __truediv__(self, rhs):
        if isinstance(rhs, Tag):
            rhs = rhs._symbol
        if isinstance(rhs, str):
            return Tag(f"{self._symbol}/{rhs}")
        return NotImplemented


class default_tags:
    parameter_quantizer = Tag("parameter")
    weight_quantizer = parameter_quantizer / "weight"
    bias_quantizer = parameter_quantizer / "bias"
    activation_quantizer = Tag("activation")
    input_quantizer = activation_quantizer / "input"
    output_quantizer = activation_quantizer / "output"


class _HasTag:
    def __init__(self, tag: Tag) -> None:
        self._tag = tag

    def __get__(self, instance, owner=None) -> bool:
        return self._tag in instance


class QuantizerMetadata:
    parameter_quantizer = _HasTag(default_tags.parameter_quantizer)
    weight_quantizer = _HasTag(default_tags.weight_quantizer)
    bias_quantizer = _HasTag(default_tags.bias_quantizer)
    input_quantizer = _HasTag(default_tags.input_quantizer)
    activation_quantizer = _HasTag(default_tags.activation_quantizer)
    output_quantizer = _HasTag(default_tags.output_quantizer)

    def __init__(self, *tags, weight_quantizer=False, bias_quantizer=False, input_quantizer=False,
                 output_quantizer=False, shape=None, **kwargs: Any) -> None:
        self._tags = set()
        self._kwargs = dict(kwargs, shape=shape)
        for tag in tags:
            self.add_tag(tag)
        for flag, tag in ((weight_quantizer, default_tags.weight_quantizer), (bias_quantizer, default_tags.bias_quantizer),
                          (input_quantizer, default_tags.input_quantizer), (output_quantizer, default_tags.output_quantizer)):
            if flag:
                self.add_tag(tag)

    def add_tag(self, tag) -> None:
        self._tags.update(Tag(tag).hierarchy())     # 'parameter/weight' also adds 'parameter'

    def __contains__(self, tag) -> bool:
        return Tag(tag) in self._tags

    def __getattr__(self, key: str) -> Any:
        kwargs = self.__dict__.get("_kwargs", {})
        if key in kwargs:
            return kwargs[key]
        raise AttributeError(key)

    @property
    def shape(self):
        return self._kwargs.get("shape")

    def is_extension(self, other: "QuantizerMetadata") -> bool:
        if not self._tags.issubset(other._tags):
            return False
        return all((k == "shape" and v is None) or other._kwargs.get(k) == v for k, v in self._kwargs.items())

    def __getstate__(self):
        return self.__dict__.copy()

    def __setstate__(self, state):
        self.__dict__.update(state)

    def __repr__(self) -> str:
        return f"QuantizerMetadata(tags={self._tags}, {', '.join(f'{k}={v}' for k, v in self._kwargs.items())})"


_MODULE_INTERNALS = frozenset(torch.nn.Module().__dict__) | {"_ffq_init_args", "quant_metadata", "_quantizer_overrides"}


@remember_init_args
class Quantizer(torch.nn.Module):
    """``forward`` = the override stack wrapped around ``quantize`` (nn/quantizer.py:413-416)."""

    def __init__(self) -> None:
        super().__init__()
        object.__setattr__(self, "_quantizer_overrides", collections.OrderedDict())
        self.quant_metadata: Optional[QuantizerMetadata] = None
        self._register_load_state_dict_pre_hook(self._materialize_on_load)

    @classmethod
    def factory(cls, *args: Any, **kwargs: Any) -> Callable[[str, "Quantizer"], "Quantizer"]:
        def make(_name: str, _current: "Quantizer") -> "Quantizer":
            return cls(*args, **kwargs)

        make.__name__ = f"{cls.__name__}_factory"
        return make

    def quantize(self, data: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def forward(self, data: torch.Tensor) -> torch.Tensor:
        return override.apply_overrides(self, self.quantize, self._quantizer_overrides)(data)

    def register_override(self, override_fn) -> override.OverrideHandle:
        handle = override.OverrideHandle(self)
        self._quantizer_overrides[handle.handle_id] = override_fn
        return handle

    def remove_override(self, override_id: int):
        return self._quantizer_overrides.pop(override_id, None)

    @property
    def overrides(self):
        yield from self._quantizer_overrides.values()

    def is_stub(self) -> bool:
        return False

    def reset_parameters(self) -> None:
        raise NotImplementedError(f"{type(self).__name__} does not implement 'reset_parameters'")

    def extra_repr(self) -> str:
        text = super().extra_repr()
        if self._quantizer_overrides:
            text += "\n(overrides): \n" + "".join(
                f"  ({i}): {fn}\n" for i, fn in enumerate(self._quantizer_overrides.values()))
        return text

    def __getstate__(self):
        if self._quantizer_overrides:
            raise RuntimeError("quantizer_overrides can not be serialized. Please remove all overrides before serialization.")
        return super().__getstate__()

    # ---- config.yaml form (serialization.py; reference nn/quantizer.py:300-339): the constructor arguments plus the
    # plain attributes; tensors travel in the safetensors file, the metadata is assigned when the quantizer is attached
    def _yaml_state(self) -> dict:
        if self._quantizer_overrides:
            raise RuntimeError("quantizer_overrides can not be serialized. Please remove all overrides before serialization.")
        return {k: v for k, v in self.__dict__.items() if k not in _MODULE_INTERNALS and not isinstance(v, torch.Tensor)}

    def _yaml_setstate(self, state: dict) -> None:
        for key, value in state.items():
            setattr(self, key, value)

    def __deepcopy__(self, memo):
        if id(self) in memo:
            return memo[id(self)]
        clone = type(self).__new__(type(self))
        memo[id(self)] = clone
        state = torch.nn.Module.__getstate__(self) if hasattr(torch.nn.Module, "__getstate__") else self.__dict__
        torch.nn.Module.__setstate__(clone, copy.deepcopy(dict(state), memo))
        return clone

    def _materialize_on_load(self, state_dict, prefix, *_unused: Any) -> None:
        """Lazy parameters take the shape of what is being loaded (nn/quantizer.py:438-463)."""
        lazy = torch.nn.parameter.UninitializedTensorMixin
        for full_name, loaded in state_dict.items():
            if not full_name.startswith(prefix):
                continue
            target = getattr(self, full_name[len(prefix):], None)
            if loaded is None or target is None:
                continue
            if isinstance(target, lazy) and not isinstance(loaded, lazy):
                with torch.no_grad():
                    target.materialize(loaded.shape)


class QuantizerStub(Quantizer):
    """Placeholder quantizer: returns its input unchanged; carries the slot's metadata."""

    def __init__(self, *tags, weight_quantizer=False, bias_quantizer=False, input_quantizer=False,
                 output_quantizer=False, shape=None, _metadata: Optional[QuantizerMetadata] = None, **kwargs: Any) -> None:
        super().__init__()
        self.quant_metadata = _metadata if _metadata is not None else QuantizerMetadata(
            *tags, weight_quantizer=weight_quantizer, bias_quantizer=bias_quantizer, input_quantizer=input_quantizer,
            output_quantizer=output_quantizer, shape=shape, **kwargs)

    def quantize(self, data: torch.Tensor) -> torch.Tensor:
        return data

    def is_stub(self) -> bool:
        return True

    def reset_parameters(self) -> None:
        pass
