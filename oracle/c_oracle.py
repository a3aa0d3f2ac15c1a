"""ctypes loader of the plain-C oracle (oracle/ffq_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "build", "libffq_oracle.so")


def load():
    if not os.path.exists(_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return ctypes.CDLL(_PATH)


def _arr(vals):
    return (ctypes.c_int64 * len(vals))(*[int(v) for v in vals])


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def quantize(lib, x, scale, tile, num_bits, offset=None):
    x = x.contiguous().float(); q = torch.empty_like(x)
    lib.ffqo_quantize(_p(x), _p(q), ctypes.c_int64(x.numel()), x.dim(), _arr(x.shape), _arr(tile), _p(scale.contiguous()),
                      _p(None if offset is None else offset.contiguous()), ctypes.c_double(num_bits))
    return q


def dequantize(lib, q, scale, tile, offset=None):
    q = q.contiguous().float(); y = torch.empty_like(q)
    lib.ffqo_dequantize(_p(q), _p(y), ctypes.c_int64(q.numel()), q.dim(), _arr(q.shape), _arr(tile), _p(scale.contiguous()),
                        _p(None if offset is None else offset.contiguous()))
    return y


def backward(lib, x, g, scale, tile, num_bits, offset=None):
    x = x.contiguous().float(); g = g.contiguous().float(); dx = torch.empty_like(x)
    nt = scale.numel()
    dscale = torch.empty(nt, dtype=torch.float64); doffset = torch.empty(nt, dtype=torch.float64) if offset is not None else None
    lib.ffqo_backward(_p(x), _p(g), _p(dx), _p(dscale), _p(doffset), ctypes.c_int64(x.numel()), ctypes.c_int64(nt), x.dim(),
                      _arr(x.shape), _arr(tile), _p(scale.contiguous()), _p(None if offset is None else offset.contiguous()),
                      ctypes.c_double(num_bits))
    return dx, dscale, doffset


def minmax(lib, x, tile, ntiles):
    x = x.contiguous().float(); mn = torch.empty(ntiles); mx = torch.empty(ntiles)
    lib.ffqo_minmax(_p(x), _p(mn), _p(mx), ctypes.c_int64(x.numel()), ctypes.c_int64(ntiles), x.dim(), _arr(x.shape), _arr(tile))
    return mn, mx


def params_for_range(lib, mn, mx, num_bits, symmetric, allow_one_sided):
    mn = mn.contiguous().float().reshape(-1); mx = mx.contiguous().float().reshape(-1)
    scale = torch.empty_like(mn); offset = torch.empty_like(mn)
    has = lib.ffqo_params_for_range(_p(mn), _p(mx), ctypes.c_int64(mn.numel()), ctypes.c_double(num_bits), int(symmetric),
                                    int(allow_one_sided), _p(scale), _p(offset))
    return scale, (offset if has else None)
