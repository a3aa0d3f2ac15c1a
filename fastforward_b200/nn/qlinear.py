"""W8A8 and W4A16 quantized linears on tcgen05 tensor cores, plugged in through the operator
dispatcher (seam 2 of SURVEY.md section 8b: ``ff.dispatcher.register("linear", predicate, kernel)``).

``install()`` registers the kernels; each predicate accepts exactly the case its kernel implements
(W8A8: int8 per-tensor activations x int8 per-channel weights; W4A16: bf16/f16 activations x
int8-stored per-group / per-channel / per-tensor weight codes) and everything else keeps taking the
reference's dequantize-then-float fallback (_gen/fallback.py:77-112)."""

from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from .. import _cabi as C
from ..dispatcher import Predicate, register
from ..quantization import granularity as G
from ..quantization.affine.function import StaticAffineQuantParams
from ..quantized_tensor import QuantizedTensor

_hook = None
_stats: Dict[str, Any] = {"calls": 0}
# bench.py's kernel census replays recorded C-ABI launches: while this list is not None the
# temporaries of every call are kept alive so that the recorded device pointers stay valid.
keepalive: Optional[list] = None


def _params(t) -> Optional[StaticAffineQuantParams]:
    if not isinstance(t, QuantizedTensor):
        return None
    p = t.quant_args()
    return p if isinstance(p, StaticAffineQuantParams) else None


def _accepts(input=None, weight=None, bias=None, output_quantizer=None, strict_quantization=None) -> bool:
    px, pw = _params(input), _params(weight)
    if px is None or pw is None or not input.is_cuda or not weight.is_cuda:
        return False
    if input.raw_data.dtype != torch.int8 or weight.raw_data.dtype != torch.int8:
        return False
    if px.num_bits > 8 or pw.num_bits > 8 or weight.dim() != 2 or input.dim() < 2:
        return False
    if not G.is_per_tensor(px.granularity):
        return False
    if not (G.is_per_channel(pw.granularity) and tuple(pw.granularity.channel_dims) == (0,)):
        return False
    if not all(isinstance(v, torch.Tensor) and v.dtype == torch.float32 for v in (px.scale, pw.scale)):
        return False
    for off in (px.offset, pw.offset):
        if off is not None and not (isinstance(off, torch.Tensor) and off.dtype == torch.float32):
            return False
    if isinstance(bias, QuantizedTensor):
        return False
    k = weight.shape[1]
    return input.shape[-1] == k and k % 16 == 0 and (px.dequantize_dtype in (torch.float32, torch.bfloat16, torch.float16))


def _attached_rowsum(t, rows: int) -> Optional[torch.Tensor]:
    rs = getattr(t, "_ffq_rowsum", None)
    if isinstance(rs, torch.Tensor) and rs.dtype == torch.int32 and rs.numel() == rows and rs.device == t.device:
        return rs
    return None


def w8a8_linear(input=None, weight=None, bias=None, output_quantizer=None, strict_quantization=None):
    px, pw = input.quant_args(), weight.quant_args()
    qx = input.raw_data
    lead = qx.shape[:-1]
    k = qx.shape[-1]
    qx2 = qx.reshape(-1, k).contiguous()
    qw = weight.raw_data.contiguous()
    m, n = qx2.shape[0], qw.shape[0]
    out_dtype = px.dequantize_dtype or torch.float32
    y = torch.empty((m, n), dtype=out_dtype, device=qx.device)
    stream = C.current_stream(qx.device)
    # row sums of the codes: produced by the fused calibration step when the codes come straight from it
    rowsum_w = _attached_rowsum(weight, n)
    if rowsum_w is None:
        rowsum_w = torch.empty(n, dtype=torch.int32, device=qx.device)
        C.check(C.lib.ffq_rowsum_i8(qw.data_ptr(), rowsum_w.data_ptr(), n, k, stream))
    rowsum_x = None
    if pw.offset is not None:
        rowsum_x = _attached_rowsum(input, m)
        if rowsum_x is None:
            rowsum_x = torch.empty(m, dtype=torch.int32, device=qx.device)
            C.check(C.lib.ffq_rowsum_i8(qx2.data_ptr(), rowsum_x.data_ptr(), m, k, stream))
    sx = px.scale.detach().reshape(-1)
    ox = None if px.offset is None else px.offset.detach().reshape(-1)
    sw = pw.scale.detach().reshape(-1).contiguous()
    ow = None if pw.offset is None else pw.offset.detach().reshape(-1).contiguous()
    b = None if bias is None else bias.detach().contiguous()
    C.check(C.lib.ffq_qlinear_w8a8(
        qx2.data_ptr(), qw.data_ptr(), y.data_ptr(), C.dtype_tag(out_dtype), m, n, k,
        sx.data_ptr(), C.ptr(ox), sw.data_ptr(), C.ptr(ow), rowsum_w.data_ptr(), C.ptr(rowsum_x),
        C.ptr(b), C.dtype_tag(b.dtype if b is not None else None), None, 0, stream))
    _stats["calls"] += 1
    if keepalive is not None:
        keepalive.append((qx2, qw, y, rowsum_w, rowsum_x, sx, ox, sw, ow, b))
    y = y.reshape(*lead, n)
    if output_quantizer is not None:
        y = output_quantizer(y)
    return y


# ------------------------------------------------------------------------------------------
# W4A16: 16-bit float activations x integer (<= 8 bit, stored as int8) per-group weights
# ------------------------------------------------------------------------------------------
_W4_BK = 64
# The fused kernel dequantizes the weight tile once per M tile; beyond ~512 rows dequantizing the weight ONCE
# (dequantize_by_tile kernel) and running the library bf16 GEMM -- the dispatcher's fallback -- is faster on B200
# (measured on 14336x4096, g=128: 46 vs 70 us at 256 rows, 83 vs 88 us at 512, 156 vs 126 us at 1024;
# tools/bench_w4a16.py).  None removes the limit.
W4A16_MAX_ROWS: Optional[int] = 512


def _w4_group(weight, pw) -> Optional[int]:
    """Elements of K sharing one weight parameter, or None if the tiling is not [1, g] / per-tensor."""
    n, k = weight.shape
    tile = pw.granularity.tile_size(weight.shape)
    tile = (n, k) if isinstance(tile, str) else tuple(tile)      # PerTensor answers "data_shape"
    if tile == (n, k):
        return k                     # per-tensor: one parameter, expanded to [N] by the wrapper
    if tile[0] != 1 or k % tile[1] != 0:
        return None
    return tile[1]


def _accepts_w4a16(input=None, weight=None, bias=None, output_quantizer=None, strict_quantization=None) -> bool:
    pw = _params(weight)
    if pw is None or not weight.is_cuda or weight.dim() != 2 or weight.raw_data.dtype != torch.int8 or pw.num_bits > 8:
        return False
    if not isinstance(input, torch.Tensor) or not input.is_cuda or input.dim() < 2:
        return False
    if isinstance(input, QuantizedTensor):
        px = _params(input)
        if px is None or _accepts(input, weight, bias, output_quantizer, strict_quantization):
            return False             # int8 x int8 belongs to the W8A8 kernel
        x_dtype = px.dequantize_dtype
    else:
        x_dtype = input.dtype
    if x_dtype not in (torch.bfloat16, torch.float16) or pw.dequantize_dtype != x_dtype:
        return False
    if not (isinstance(pw.scale, torch.Tensor) and pw.scale.dtype == torch.float32):
        return False
    if pw.offset is not None and not (isinstance(pw.offset, torch.Tensor) and pw.offset.dtype == torch.float32):
        return False
    if isinstance(bias, QuantizedTensor):
        return False
    k = weight.shape[1]
    if W4A16_MAX_ROWS is not None and k > 0 and input.numel() // k > W4A16_MAX_ROWS:
        return False
    group = _w4_group(weight, pw)
    return input.shape[-1] == k and k % _W4_BK == 0 and group is not None and group % _W4_BK == 0


def w4a16_linear(input=None, weight=None, bias=None, output_quantizer=None, strict_quantization=None):
    pw = weight.quant_args()
    x = input.dequantize() if isinstance(input, QuantizedTensor) else input
    lead = x.shape[:-1]
    k = x.shape[-1]
    x2 = x.detach().reshape(-1, k).contiguous()
    qw = weight.raw_data.contiguous()
    m, n = x2.shape[0], qw.shape[0]
    group = _w4_group(weight, pw)
    groups = k // group
    sw = pw.scale.detach().reshape(-1)
    ow = None if pw.offset is None else pw.offset.detach().reshape(-1)
    if sw.numel() == 1 and n * groups != 1:            # per-tensor weight parameters
        sw = sw.expand(n * groups)
        ow = None if ow is None else ow.expand(n * groups)
    sw = sw.contiguous()
    ow = None if ow is None else ow.contiguous()
    b = None if bias is None else bias.detach().contiguous()
    y = torch.empty((m, n), dtype=x2.dtype, device=x2.device)
    C.check(C.lib.ffq_qlinear_w4a16(
        x2.data_ptr(), C.dtype_tag(x2.dtype), qw.data_ptr(), y.data_ptr(), m, n, k,
        sw.data_ptr(), C.ptr(ow), group, C.ptr(b), C.dtype_tag(b.dtype if b is not None else None),
        C.current_stream(x2.device)))
    _stats["calls_w4a16"] = _stats.get("calls_w4a16", 0) + 1
    if keepalive is not None:
        keepalive.append((x2, qw, y, sw, ow, b))
    y = y.reshape(*lead, n)
    if output_quantizer is not None:
        y = output_quantizer(y)
    return y


_hook_w4 = None


def install() -> None:
    """Register the kernels (idempotent)."""
    global _hook, _hook_w4
    if _hook is None:
        _hook = register("linear", Predicate(_accepts), w8a8_linear)
    if _hook_w4 is None:
        _hook_w4 = register("linear", Predicate(_accepts_w4a16), w4a16_linear)


def uninstall() -> None:
    global _hook, _hook_w4
    if _hook is not None:
        _hook.remove()
        _hook = None
    if _hook_w4 is not None:
        _hook_w4.remove()
        _hook_w4 = None


def stats() -> Dict[str, Any]:
    return dict(_stats)
