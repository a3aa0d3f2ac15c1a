"""ctypes binding of the C ABI declared in include/ffq_b200.h.

The shared library is the product's only compute path.  There is no CPU or PyTorch fallback:
if ``lib/libffq_b200.so`` is missing this module raises at import time, and every op wrapper
refuses non-CUDA tensors.
"""

from __future__ import annotations

import ctypes
import functools
import os
from typing import Optional, Sequence

import torch  # imported first on purpose: it loads libcudart.so.12, which the library links against

from .exceptions import QuantizationError  # noqa: F401  (re-exported for callers)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FFQ_LIB_PATH") or os.path.join(_HERE, "lib", "libffq_b200.so")   # override: A/B of two builds

FFQ_MAX_RANK = 8

# ffq_status_t
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_BITWIDTH, ERR_WORKSPACE = range(6)
# ffq_ws_kind_t
WS_QUANTIZE_BWD, WS_MINMAX, WS_PARAMS_FOR_RANGE, WS_DYNAMIC_QUANTIZE = range(4)

# ffq_dtype_t
DT_NONE = 255
_DTYPES = {
    torch.float32: 0,
    torch.float16: 1,
    torch.bfloat16: 2,
    torch.float64: 3,
    torch.int8: 4,
    torch.int16: 5,
    torch.int32: 6,
    torch.uint8: 7,
    torch.int64: 8,
}


def dtype_tag(dtype: Optional[torch.dtype]) -> int:
    if dtype is None:
        return DT_NONE
    try:
        return _DTYPES[dtype]
    except KeyError:
        raise NotImplementedError(f"fastforward_b200: dtype {dtype} is not supported by the B200 backend") from None


class Layout(ctypes.Structure):
    """ffq_layout_t"""

    _fields_ = [
        ("rank", ctypes.c_int32),
        ("dims", ctypes.c_int64 * FFQ_MAX_RANK),
        ("tile", ctypes.c_int64 * FFQ_MAX_RANK),
    ]


class Requant(ctypes.Structure):
    """ffq_requant_t: the output quantizer fused into the W8A8 epilogue"""

    _fields_ = [
        ("scale", ctypes.c_void_p),
        ("offset", ctypes.c_void_p),
        ("num_bits", ctypes.c_double),
        ("codes", ctypes.c_void_p),
        ("rowsum", ctypes.c_void_p),
    ]


ABI_VERSION = 2


@functools.lru_cache(maxsize=4096)
def make_layout(shape: tuple, tile: tuple) -> Layout:
    if len(shape) != len(tile):
        raise ValueError(
            f"Input dimensionality must match tile_size dimensionality got {len(shape)} and {len(tile)}"
        )
    if len(shape) > FFQ_MAX_RANK:
        raise NotImplementedError(f"fastforward_b200 supports tensors of rank <= {FFQ_MAX_RANK}")
    bad = [i for i, (d, t) in enumerate(zip(shape, tile)) if t > 0 and d % t != 0]
    if bad or any(t <= 0 for t in tile):
        raise ValueError(
            "Each dimension of tile_size must divide the corresponding input dimension. Got "
            + ", ".join(f"{shape[i]} and {tile[i]} for dimension {i}" for i in (bad or range(len(tile))))
            + "."
        )
    lay = Layout()
    lay.rank = len(shape)
    ntiles = 1
    for i, (d, t) in enumerate(zip(shape, tile)):
        lay.dims[i] = d
        lay.tile[i] = t
        ntiles *= d // t
    lay.ref = ctypes.byref(lay)      # cached: the object lives in the lru_cache for the life of the process
    lay.num_tiles = ntiles
    lay.ws = {}                      # (kind, dtype) -> workspace bytes
    return lay


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"fastforward_b200: CUDA library not found at {LIB_PATH}. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C fastforward_b200/csrc`). "
            "There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_size_t
    lp = ctypes.POINTER(Layout)
    sig = {
        "ffq_abi_version": (ctypes.c_int, []),
        "ffq_last_error": (ctypes.c_char_p, []),
        "ffq_launch_count": (ctypes.c_uint64, []),
        "ffq_workspace_bytes": (sz, [i32, lp, i32]),
        "ffq_num_tiles": (i64, [lp]),
        "ffq_quantize": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, lp, dbl, vp]),
        "ffq_dequantize": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, lp, vp]),
        "ffq_fakequant_fwd": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, lp, dbl, vp]),
        "ffq_quantize_bwd": (i32, [vp, i32, vp, i32, vp, vp, vp, vp, i32, vp, i32, lp, dbl, vp, sz, vp]),
        "ffq_minmax": (i32, [vp, i32, vp, vp, vp, vp, i32, vp, lp, vp, sz, vp]),
        "ffq_params_for_range": (i32, [vp, vp, i32, i64, dbl, i32, i32, i32, vp, i32, vp, i32, vp, sz, vp]),
        "ffq_dynamic_quantize": (i32, [vp, i32, vp, i32, vp, vp, lp, dbl, i32, i32, vp, sz, vp]),
        "ffq_qlinear_w8a8": (i32, [vp, vp, vp, i32, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, i32, ctypes.POINTER(Requant), vp]),
        "ffq_rowsum_i8": (i32, [vp, vp, i64, i64, vp]),
        "ffq_fakequant_fwd_bwd_host": (i32, [vp, vp, i32, vp, vp, vp, vp, vp, vp, lp, dbl, i32]),
        "ffq_selftest_shared_div": (i32, [ctypes.c_uint64, ctypes.c_uint32, vp, vp]),
        "ffq_qlinear_w4a16": (i32, [vp, i32, vp, vp, i64, i64, i64, vp, vp, i64, vp, i32, vp]),
        "ffq_grid_mse": (i32, [vp, i32, vp, vp, i32, vp, lp, dbl, vp, sz, vp]),
        "ffq_grid_mse_workspace_bytes": (sz, [lp, i32, i32]),
        "ffq_calibrate_quantize": (i32, [vp, i32, vp, vp, vp, i32, vp, vp, vp, i64, vp, vp, i32, lp, dbl, i32, i32, vp, sz, vp]),
        "ffq_calibrate_quantize_mode": (i32, [lp, i32]),
        "ffq_calibrate_quantize_workspace_bytes": (sz, []),
        "ffq_calibrate_fakequant": (i32, [vp, i32, vp, vp, vp, i32, vp, vp, vp, lp, dbl, i32, i32, i32, vp, sz, vp]),
        "ffq_params_for_ranges_batched": (i32, [vp, vp, i32, vp, i64, vp]),
        "ffq_params_for_ranges_encode": (None, [dbl, i32, i32, ctypes.POINTER(ctypes.c_int64)]),
        "ffq_debug_gemm_profile": (None, [vp]),
        "ffq_calibrate_fakequant_batched": (i32, [vp, vp, i64, i64, i32, i64, dbl, i32, i32, i32, vp, sz, vp]),
        "ffq_gptq_block": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, i64, i64, vp, i32, vp, i32, vp, i64, i64, i64, dbl, i32, vp]),
        "ffq_lpbq_encode": (i32, [vp, i64, i64, i32, i32, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.ffq_abi_version() != ABI_VERSION:
        raise ImportError(f"fastforward_b200: ABI version mismatch ({lib.ffq_abi_version()} != {ABI_VERSION}); rebuild "
                          "the library (make -C fastforward_b200/csrc)")
    return lib


lib = _load()
EXPORTED = (
    "ffq_abi_version ffq_last_error ffq_launch_count ffq_workspace_bytes ffq_num_tiles ffq_quantize "
    "ffq_dequantize ffq_fakequant_fwd ffq_quantize_bwd ffq_minmax ffq_params_for_range "
    "ffq_dynamic_quantize ffq_qlinear_w8a8 ffq_rowsum_i8 ffq_fakequant_fwd_bwd_host ffq_selftest_shared_div ffq_grid_mse ffq_grid_mse_workspace_bytes ffq_qlinear_w4a16 "
    "ffq_calibrate_quantize ffq_calibrate_quantize_mode ffq_calibrate_quantize_workspace_bytes ffq_gptq_block "
    "ffq_params_for_ranges_batched ffq_params_for_ranges_encode ffq_calibrate_fakequant ffq_debug_gemm_profile ffq_calibrate_fakequant_batched ffq_lpbq_encode"
).split()


def last_error() -> str:
    msg = lib.ffq_last_error()
    return msg.decode() if msg else ""


def check(rc: int) -> None:
    """Map an ffq_status_t to the exception type the reference raises for that condition
    (SURVEY.md section 8b 'Error conventions')."""
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f"fastforward_b200: {msg} (status {rc})")


def launch_count() -> int:
    return int(lib.ffq_launch_count())


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def current_stream(device: torch.device) -> int:
    if _raw_stream is not None:
        return _raw_stream(device.index if device.index is not None else torch.cuda.current_device())
    return torch.cuda.current_stream(device).cuda_stream


_get_device = getattr(torch._C, "_cuda_getDevice", None)


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_of(device: torch.device):
    """Context manager that makes ``device`` current for the C-ABI calls inside it -- the library launches on the
    stream it is given but reads per-device state (SM count, opt-in shared-memory attributes, occupancy) from the
    current device.  Free when the device is already current (the usual one-process-per-GPU case)."""
    idx = device.index
    if idx is None or _get_device is None or idx == _get_device():
        return _NO_GUARD
    return torch.cuda.device(idx)


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"fastforward_b200: '{what}' is on {t.device}; this backend only runs on CUDA (sm_100a) "
            "and has no CPU fallback."
        )


def workspace_bytes(kind: int, layout: Layout, dtype: torch.dtype) -> int:
    key = (kind, dtype)
    cache = getattr(layout, "ws", None)
    if cache is not None and key in cache:
        return cache[key]
    n = int(lib.ffq_workspace_bytes(kind, ctypes.byref(layout), dtype_tag(dtype)))
    if cache is not None:
        cache[key] = n
    return n


_scratch: dict = {}


def scratch(device: torch.device, nbytes: int) -> torch.Tensor:
    """Small persistent per-(device, stream) scratch buffer for parameter-sized kernels."""
    key = (device.index, current_stream(device))
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 8192), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


_barrier_ws: dict = {}


def barrier_workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Persistent ZERO-initialised per-(device, stream) workspace for kernels with a grid barrier
    (ffq_calibrate_quantize): the kernel leaves the barrier words zeroed for its next launch."""
    key = (device.index, current_stream(device))
    buf = _barrier_ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _barrier_ws[key] = buf
    return buf
