"""``QuantizedTensor``: a ``torch.Tensor`` subclass holding raw integer codes plus the context
needed to dequantize them (reference: quantized_tensor.py:290-598).

Behaviour kept from the reference:
  * ``dequantize()`` routes to the context's quantization function -> the CUDA dequantize kernel;
  * ``.to(device)`` moves the parameters too; ``.to(dtype)`` / ``.float()`` ... dequantize first;
  * ``clone/detach/contiguous/cpu/cuda/deepcopy/pickle`` keep the context;
  * a fixed set of pure queries bypasses dispatch; every other torch function goes through
    ``dispatcher.dispatch(func.__name__, ...)`` and otherwise falls back to implicit
    dequantization, which raises ``QuantizationError`` under strict quantization;
  * in-place tensor methods are not implemented unless a kernel is registered for them.
"""

from __future__ import annotations

import copy
import functools
import warnings
from typing import Any, Callable

import torch
from torch._C import DisableTorchFunctionSubclass

from . import flags
from .dispatcher import DispatcherPriority, dispatch, register
from .exceptions import QuantizationError


def _to_dtype(dtype: torch.dtype, qtensor: "QuantizedTensor") -> torch.Tensor:
    return qtensor.dequantize().to(dtype)


for _name, _dtype in (
    ("double", torch.double), ("float", torch.float), ("half", torch.half), ("bfloat16", torch.bfloat16),
    ("long", torch.int64), ("int", torch.int32), ("short", torch.int16), ("char", torch.int8),
    ("bool", torch.bool), ("byte", torch.uint8), ("cdouble", torch.complex128), ("cfloat", torch.complex64),
):
    register(_name, None, functools.partial(_to_dtype, _dtype))


def _not_implemented(name: str, inplace: bool) -> Callable[..., Any]:
    kind = "The in-place operation" if inplace else "The operation"
    msg = (
        f"{kind} '{name}' is not implemented for QuantizedTensor. A user implementation of {name} can be "
        "registered through the QuantizedTensor dispatcher system (fastforward_b200.dispatcher.register)."
    )

    def raiser(*_a: Any, **_k: Any):
        raise NotImplementedError(msg)

    raiser.__name__ = f"{name}_not_implemented"
    return raiser


for _name in ("__getitem__", "__reversed__", "__setitem__"):
    register(_name, None, _not_implemented(_name, False), DispatcherPriority.NOT_IMPLEMENTED_FALLBACK)
for _name in dir(torch.Tensor):
    if _name.endswith("_") and not _name.endswith("__") and callable(getattr(torch.Tensor, _name, None)):
        register(_name, None, _not_implemented(_name, True), DispatcherPriority.NOT_IMPLEMENTED_FALLBACK)

# Pure queries / bookkeeping that must not dequantize or dispatch.
_NO_DISPATCH_NAMES = (
    "size dim ndimension numel nelement element_size stride storage_offset is_contiguous is_floating_point "
    "is_complex is_signed is_inference is_pinned is_shared is_set_to data_ptr get_device type_as "
    "requires_grad_ retain_grad register_hook backward untyped_storage has_names "
    "__len__ __hash__ __format__ __dlpack_device__ __sizeof__ _is_view is_same_size is_nonzero "
    "dim_order is_neg is_conj is_quantized is_coalesced is_sparse_csr _base"
).split()
_NO_DISPATCH = set()
for _name in _NO_DISPATCH_NAMES:
    for _owner in (torch.Tensor, torch):
        _f = getattr(_owner, _name, None)
        if callable(_f):
            _NO_DISPATCH.add(_f)
for _name in (
    "shape dtype device grad grad_fn requires_grad is_leaf is_cuda is_cpu is_meta is_sparse layout names ndim "
    "output_nr data _version retains_grad is_mkldnn is_xpu is_mps is_quantized"
).split():
    _prop = getattr(torch.Tensor, _name, None)
    for _acc in ("__get__", "__set__", "__delete__"):
        _f = getattr(_prop, _acc, None)
        if _f is not None:
            _NO_DISPATCH.add(_f)


def _rebuild_quantized_tensor(raw_data: torch.Tensor, context: Any) -> "QuantizedTensor":
    return QuantizedTensor(raw_data, context)


class QuantizedTensor(torch.Tensor):
    @staticmethod
    def __new__(cls, data: torch.Tensor, *_a: Any, **_k: Any) -> "QuantizedTensor":
        return data.as_subclass(cls)

    def __init__(self, data: torch.Tensor, quantization_context: Any) -> None:
        super().__init__()
        self._quantization_context = quantization_context

    # ---- accessors ---------------------------------------------------------------------
    @property
    def raw_data(self) -> torch.Tensor:
        return self.as_subclass(torch.Tensor)

    def int_repr(self) -> torch.Tensor:
        return self.raw_data

    # In-place operators are declined (quantized_tensor.py:488-507): the result of an arithmetic operator does not lie
    # on the quantization grid, so ``qt += x`` must not write into the codes; Python then evaluates ``qt = qt + x``.
    def _declined(self, *args: Any, **kwargs: Any):
        return NotImplemented

    __iadd__ = __isub__ = __imul__ = __imatmul__ = __itruediv__ = __ifloordiv__ = __imod__ = _declined
    __ilshift__ = __irshift__ = __iand__ = __ixor__ = __ior__ = __ipow__ = _declined

    def quant_args(self):
        return self._quantization_context.quantization_params

    @property
    def quantization_context(self):
        return self._quantization_context

    @property
    def quant_func(self):
        return self._quantization_context.quantization_fn

    def dequantize(self) -> torch.Tensor:
        ctx = self._quantization_context
        return ctx.quantization_fn.dequantize(self.raw_data, ctx.quantization_params)

    @property
    def is_quantized(self) -> bool:  # type: ignore[override]
        warnings.warn(
            "QuantizedTensor.is_quantized refers to PyTorch's native quantized tensors and is False; "
            "use isinstance(x, QuantizedTensor)."
        )
        return False

    # ---- movement / copies ----------------------------------------------------------------
    def to(self, *args: Any, **kwargs: Any):  # type: ignore[override]
        if (args and isinstance(args[0], torch.Tensor)) or "other" in kwargs:
            raise ValueError(f"{type(self).__name__}.to(other: Tensor, ...) is not supported")
        device, dtype, non_blocking, memory_format = torch._C._nn._parse_to(*args, **kwargs)
        if dtype is not None:
            return self.dequantize().to(device=device, dtype=dtype, non_blocking=non_blocking,
                                        memory_format=memory_format)
        with DisableTorchFunctionSubclass():
            moved = super().to(device=device, non_blocking=non_blocking, memory_format=memory_format)
        return type(self)(moved, self._quantization_context.to(device))

    def cuda(self, device=None, non_blocking: bool = False):  # type: ignore[override]
        return self.to(device=device or "cuda", non_blocking=non_blocking)

    def cpu(self):  # type: ignore[override]
        return self.to("cpu")

    def clone(self):  # type: ignore[override]
        ctx = self._quantization_context.clone_parameters()
        with DisableTorchFunctionSubclass():
            data = super().clone()
        return ctx.attach(data)

    def detach(self):  # type: ignore[override]
        ctx = self._quantization_context.detach_parameters()
        with DisableTorchFunctionSubclass():
            data = super().detach()
        return ctx.attach(data)

    def contiguous(self, memory_format=torch.contiguous_format):  # type: ignore[override]
        ctx = self._quantization_context.contiguous_parameters()
        with DisableTorchFunctionSubclass():
            data = super().contiguous(memory_format=memory_format)
        if ctx is self._quantization_context and data.data_ptr() == self.data_ptr():
            return self
        return ctx.attach(data)

    def view(self, *shape: Any):  # same-shape views are always legal
        target = shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape
        if isinstance(target, torch.dtype):
            raise NotImplementedError("view(dtype) is not implemented for QuantizedTensor")
        if tuple(target) == tuple(self.shape):
            return self
        return self.__torch_function__(torch.Tensor.view, (type(self),), (self,) + tuple(shape))

    def view_as(self, other: torch.Tensor):
        return self.view(other.shape)

    def __deepcopy__(self, memo: dict):
        if not self.is_leaf:
            raise RuntimeError(
                "Only Tensors created explicitly by the user (graph leaves) support the deepcopy protocol at the moment"
            )
        return type(self)(copy.deepcopy(self.raw_data.detach(), memo), copy.deepcopy(self._quantization_context, memo))

    def __reduce_ex__(self, proto: int):
        return _rebuild_quantized_tensor, (self.raw_data.detach(), self._quantization_context)

    # ---- dispatch -------------------------------------------------------------------------------
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        with DisableTorchFunctionSubclass():
            if func in _NO_DISPATCH:
                return func(*args, **kwargs)
            name = getattr(func, "__name__", None)
            if name:
                kernel = dispatch(name, *args, **kwargs)
                if kernel is not None:
                    return kernel(*args, **kwargs)
            return _dequantization_fallback(func, *args, **kwargs)

    def __repr__(self, **kwargs: Any) -> str:  # type: ignore[override]
        with torch._C.DisableTorchFunction():
            return f"QuantizedTensor({self.raw_data!r}, {self._quantization_context.quantization_params!r})"


def _dequantization_fallback(func: Callable[..., Any], *args: Any, **kwargs: Any) -> Any:
    """Implicitly dequantize every QuantizedTensor argument and call the torch function
    (quantized_tensor.py:548-563).  Forbidden under strict quantization."""
    if flags.get_strict_quantization():
        raise QuantizationError(
            f"'{getattr(func, '__name__', func)}' has no quantized implementation and implicit dequantization is "
            "disallowed because strict_quantization is enabled. Dequantize explicitly, register a kernel "
            "through fastforward_b200.dispatcher.register, or disable strict quantization."
        )
    from torch.utils._pytree import tree_map

    def deq(v: Any) -> Any:
        return v.dequantize() if isinstance(v, QuantizedTensor) else v

    return func(*tree_map(deq, args), **tree_map(deq, kwargs))
