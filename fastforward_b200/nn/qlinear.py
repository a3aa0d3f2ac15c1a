"""W8A8 and W4A16 quantized linears (and the int8 ``matmul`` / ``mm`` / ``bmm``) on tcgen05 tensor cores, plugged in
through the operator dispatcher (seam 2 of SURVEY.md section 8b: ``ff.dispatcher.register("linear", predicate, kernel)``).

``install()`` registers the kernels with this package's dispatcher; ``build(host)`` makes the same predicates and
kernels against ANOTHER host package's classes -- ``fastforward_b200.plugin`` uses it to register them with the
unmodified reference (its ``QuantizedTensor`` / ``StaticAffineQuantParams`` differ from ours only by identity).

Each predicate accepts exactly the case its kernel implements (W8A8: int8 per-tensor activations x int8 per-channel
weights; W4A16: bf16/f16 activations x int8-stored per-group / per-channel / per-tensor weight codes) and everything
else keeps taking the reference's dequantize-then-float fallback (_gen/fallback.py:77-112).

The output quantizer (fallback.py:109-110) is fused into the W8A8 epilogue when it is a calibrated per-tensor int8
``LinearQuantizer`` with nothing overriding it: the kernel then writes the int8 codes (and their row sums, which the
next W8A8 linear needs) instead of a bf16 tensor that would be read back and quantized by a second kernel."""

from __future__ import annotations

import ctypes
import types
from typing import Any, Dict, Optional

import torch

from .. import _cabi as C

_stats: Dict[str, Any] = {"calls": 0}
# bench.py's kernel census replays recorded C-ABI launches: while this list is not None the
# temporaries of every call are kept alive so that the recorded device pointers stay valid.
keepalive: Optional[list] = None

_W4_BK = 64
# Rows above which the weight-only kernel hands over to dequantize-once + library GEMM (None: never).  The kernel shares
# each dequantized weight tile between two 256-row activation tiles, which keeps it ahead of that route at every size.
W4A16_MAX_ROWS: Optional[int] = None


def _own_host():
    from .. import flags
    from ..exceptions import QuantizationError
    from ..nn.linear_quantizer import LinearQuantizer
    from ..quantization import granularity as G
    from ..quantization.affine.function import AffineQuantizationFunction, StaticAffineQuantParams
    from ..quantization.function import QuantizationContext
    from ..quantized_tensor import QuantizedTensor

    return types.SimpleNamespace(
        QuantizedTensor=QuantizedTensor, StaticAffineQuantParams=StaticAffineQuantParams, granularity=G,
        LinearQuantizer=LinearQuantizer, QuantizationContext=QuantizationContext,
        AffineQuantizationFunction=AffineQuantizationFunction, QuantizationError=QuantizationError, flags=flags)


def reference_host(ff):
    """The classes of an installed, unmodified ``fastforward`` package."""
    return types.SimpleNamespace(
        QuantizedTensor=ff.QuantizedTensor, StaticAffineQuantParams=ff.quantization.affine.StaticAffineQuantParams,
        granularity=ff.quantization.granularity, LinearQuantizer=ff.nn.LinearQuantizer,
        QuantizationContext=ff.quantization.function.QuantizationContext,
        AffineQuantizationFunction=ff.quantization.affine.AffineQuantizationFunction,
        QuantizationError=ff.exceptions.QuantizationError, flags=ff.flags)


def _attached_rowsum(t, rows: int) -> Optional[torch.Tensor]:
    rs = getattr(t, "_ffq_rowsum", None)
    if isinstance(rs, torch.Tensor) and rs.dtype == torch.int32 and rs.numel() == rows and rs.device == t.device:
        return rs
    return None


def _f32(t) -> bool:
    return isinstance(t, torch.Tensor) and t.dtype == torch.float32


def build(host) -> types.SimpleNamespace:
    """Predicates and kernels against one host package's classes."""
    QT, G = host.QuantizedTensor, host.granularity

    def params(t):
        if not isinstance(t, QT):
            return None
        p = t.quant_args()
        return p if isinstance(p, host.StaticAffineQuantParams) else None

    # --------------------------------------------------------------------------------------------------------
    # W8A8
    # --------------------------------------------------------------------------------------------------------
    def accepts_w8a8(input=None, weight=None, bias=None, output_quantizer=None, strict_quantization=None) -> bool:
        # the predicate reads a dozen attributes of QuantizedTensors: each would be a __torch_function__ round trip
        with torch._C.DisableTorchFunctionSubclass():
            return _accepts_w8a8(input, weight, bias, output_quantizer, strict_quantization)

    def _accepts_w8a8(input, weight, bias, output_quantizer, strict_quantization) -> bool:
        px, pw = params(input), params(weight)
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
        if not (_f32(px.scale) and _f32(pw.scale)):
            return False
        if any(o is not None and not _f32(o) for o in (px.offset, pw.offset)):
            return False
        if isinstance(bias, QT):
            return False
        if strict_quantization and output_quantizer is None:
            return False                 # the fallback raises QuantizationError for this (fallback.py:88-90)
        if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad
                                           for t in (input, weight, bias, px.scale, pw.scale, px.offset, pw.offset)):
            return False                 # no backward here: the fallback's float path carries the gradients
        k = weight.shape[1]
        return input.shape[-1] == k and k % 16 == 0 and (px.dequantize_dtype in (torch.float32, torch.bfloat16, torch.float16))

    def _fusable_output_quantizer(oq, out_dtype) -> bool:
        """A calibrated per-tensor int8 LinearQuantizer with nothing between us and its own quantize."""
        if type(oq) is not host.LinearQuantizer or getattr(oq, "quantized_dtype", None) != torch.int8:
            return False
        if oq.num_bits > 8 or not G.is_per_tensor(oq.granularity) or getattr(oq, "_quantizer_overrides", None):
            return False
        if oq.has_uninitialized_params or host.flags.get_export_mode() or torch.is_grad_enabled():
            return False
        if not (_f32(oq.scale) and oq.scale.numel() == 1 and oq.scale.is_cuda):
            return False
        return oq.offset is None or (_f32(oq.offset) and oq.offset.numel() == 1)

    def w8a8_linear(input=None, weight=None, bias=None, output_quantizer=None, strict_quantization=None):
        px, pw = input.quant_args(), weight.quant_args()
        qx = input.raw_data
        lead = qx.shape[:-1]
        k = qx.shape[-1]
        qx2 = qx.reshape(-1, k).contiguous()
        qw = weight.raw_data.contiguous()
        m, n = qx2.shape[0], qw.shape[0]
        out_dtype = px.dequantize_dtype or torch.float32
        dev = qx.device
        with C.device_of(dev):
            stream = C.current_stream(dev)
            # row sums of the codes: produced by the fused calibration step when the codes come straight from it
            rowsum_w = _attached_rowsum(weight, n)
            if rowsum_w is None:
                rowsum_w = torch.empty(n, dtype=torch.int32, device=dev)
                C.check(C.lib.ffq_rowsum_i8(qw.data_ptr(), rowsum_w.data_ptr(), n, k, stream))
            rowsum_x = None
            if pw.offset is not None:
                rowsum_x = _attached_rowsum(input, m)
                if rowsum_x is None:
                    rowsum_x = torch.empty(m, dtype=torch.int32, device=dev)
                    C.check(C.lib.ffq_rowsum_i8(qx2.data_ptr(), rowsum_x.data_ptr(), m, k, stream))
            # only the addresses are needed: a contiguous tensor is used as it is (detach / reshape / contiguous are
            # a microsecond of host time each, five to ten of them per linear)
            sx, ox, sw, ow, b = (None if t is None else (t if t.is_contiguous() else t.detach().contiguous())
                                 for t in (px.scale, px.offset, pw.scale, pw.offset, bias))
            fuse = output_quantizer is not None and _fusable_output_quantizer(output_quantizer, out_dtype)
            if fuse:
                oq = output_quantizer
                codes = torch.empty((m, n), dtype=torch.int8, device=dev)
                rs_out = torch.zeros(m, dtype=torch.int32, device=dev)
                rq = C.Requant(oq.scale.data_ptr(), C.ptr(oq.offset), float(oq.num_bits), codes.data_ptr(), rs_out.data_ptr())
                y = None
            else:
                y = torch.empty((m, n), dtype=out_dtype, device=dev)
                rq = None
            C.check(C.lib.ffq_qlinear_w8a8(
                qx2.data_ptr(), qw.data_ptr(), C.ptr(y), C.dtype_tag(out_dtype), m, n, k,
                sx.data_ptr(), C.ptr(ox), sw.data_ptr(), C.ptr(ow), rowsum_w.data_ptr(), C.ptr(rowsum_x),
                C.ptr(b), C.dtype_tag(b.dtype if b is not None else None),
                ctypes.byref(rq) if rq is not None else None, stream))
        _stats["calls"] += 1
        if keepalive is not None:
            keepalive.append((qx2, qw, y, rowsum_w, rowsum_x, sx, ox, sw, ow, b))
        if fuse:
            _stats["calls_requant_fused"] = _stats.get("calls_requant_fused", 0) + 1
            p = oq.quantization_parameters()
            p = p.with_changes(dequantize_dtype=p.dequantize_dtype or out_dtype)
            out = QT(codes.reshape(*lead, n), host.QuantizationContext(host.AffineQuantizationFunction, p))
            out._ffq_rowsum = rs_out
            return out
        y = y.reshape(*lead, n)
        if output_quantizer is not None:
            y = output_quantizer(y)
        return y

    # --------------------------------------------------------------------------------------------------------
    # W4A16: 16-bit float activations x integer (<= 8 bit, stored as int8) per-group weights
    # --------------------------------------------------------------------------------------------------------
    def w4_group(weight, pw) -> Optional[int]:
        """Elements of K sharing one weight parameter, or None if the tiling is not [1, g] / per-tensor."""
        n, k = weight.shape
        tile = pw.granularity.tile_size(weight.shape)
        tile = (n, k) if isinstance(tile, str) else tuple(tile)      # PerTensor answers "data_shape"
        if tile == (n, k):
            return k                     # per-tensor: one parameter, expanded to [N] by the wrapper
        if tile[0] != 1 or k % tile[1] != 0:
            return None
        return tile[1]

    def accepts_w4a16(input=None, weight=None, bias=None, output_quantizer=None, strict_quantization=None) -> bool:
        pw = params(weight)
        if pw is None or not weight.is_cuda or weight.dim() != 2 or weight.raw_data.dtype != torch.int8 or pw.num_bits > 8:
            return False
        if not isinstance(input, torch.Tensor) or not input.is_cuda or input.dim() < 2:
            return False
        if strict_quantization and (output_quantizer is None or not isinstance(input, QT)):
            return False                 # the fallback raises QuantizationError for these (fallback.py:88-96)
        if isinstance(input, QT):
            px = params(input)
            if px is None or accepts_w8a8(input, weight, bias, output_quantizer, strict_quantization):
                return False             # int8 x int8 belongs to the W8A8 kernel
            x_dtype = px.dequantize_dtype
        else:
            x_dtype = input.dtype
        if x_dtype not in (torch.bfloat16, torch.float16) or pw.dequantize_dtype != x_dtype:
            return False
        if not _f32(pw.scale) or (pw.offset is not None and not _f32(pw.offset)):
            return False
        if isinstance(bias, QT):
            return False
        if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad
                                           for t in (input, weight, bias, pw.scale, pw.offset)):
            return False                 # the kernel has no backward: the fallback (dequantize + F.linear) propagates
        k = weight.shape[1]
        if W4A16_MAX_ROWS is not None and k > 0 and input.numel() // k > W4A16_MAX_ROWS:
            return False
        group = w4_group(weight, pw)
        return input.shape[-1] == k and k % _W4_BK == 0 and group is not None and group % _W4_BK == 0

    def w4a16_linear(input=None, weight=None, bias=None, output_quantizer=None, strict_quantization=None):
        pw = weight.quant_args()
        x = input.dequantize() if isinstance(input, QT) else input
        lead = x.shape[:-1]
        k = x.shape[-1]
        x2 = x.detach().reshape(-1, k).contiguous()
        qw = weight.raw_data.contiguous()
        m, n = x2.shape[0], qw.shape[0]
        group = w4_group(weight, pw)
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
        with C.device_of(x2.device):
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

    # --------------------------------------------------------------------------------------------------------
    # int8 matmul / mm / bmm (the attention matmuls of the tutorial Llama, _gen/operators.py:654-735): both operands
    # per-tensor int8 QuantizedTensors.  other^T is the GEMM's [N, K] operand: a K-contiguous `other` (k.transpose(-1,-2)
    # of a contiguous k -- the QK^T case) is used as it is, an N-contiguous one (attn @ v) is transposed once.
    # --------------------------------------------------------------------------------------------------------
    def _mm_operands_ok(a, b) -> bool:
        pa, pb = params(a), params(b)
        if pa is None or pb is None or not a.is_cuda or not b.is_cuda:
            return False
        if a.raw_data.dtype != torch.int8 or b.raw_data.dtype != torch.int8 or pa.num_bits > 8 or pb.num_bits > 8:
            return False
        if not (G.is_per_tensor(pa.granularity) and G.is_per_tensor(pb.granularity)):
            return False
        if not (_f32(pa.scale) and _f32(pb.scale)) or any(o is not None and not _f32(o) for o in (pa.offset, pb.offset)):
            return False
        if pa.dequantize_dtype not in (torch.float32, torch.bfloat16, torch.float16):
            return False
        if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad
                                           for t in (a, b, pa.scale, pb.scale, pa.offset, pb.offset)):
            return False
        return a.dim() >= 2 and b.dim() >= 2 and a.shape[-1] == b.shape[-2] and a.shape[-1] % 16 == 0

    def accepts_matmul(input=None, other=None, output_quantizer=None, strict_quantization=None, **kw) -> bool:
        if kw or (strict_quantization and output_quantizer is None):
            return False
        if not _mm_operands_ok(input, other):
            return False
        if input.dim() == 2 and other.dim() == 2:
            return True
        # batched: identical leading dims (no broadcasting), handled one [M,K] x [K,N] problem at a time
        return input.dim() == other.dim() and tuple(input.shape[:-2]) == tuple(other.shape[:-2])

    def accepts_mm(input=None, mat2=None, output_quantizer=None, strict_quantization=None, **kw) -> bool:
        return not kw and isinstance(input, torch.Tensor) and isinstance(mat2, torch.Tensor) and input.dim() == 2 and \
            mat2.dim() == 2 and accepts_matmul(input, mat2, output_quantizer, strict_quantization)

    def accepts_bmm(input=None, mat2=None, output_quantizer=None, strict_quantization=None, **kw) -> bool:
        return not kw and isinstance(input, torch.Tensor) and isinstance(mat2, torch.Tensor) and input.dim() == 3 and \
            mat2.dim() == 3 and accepts_matmul(input, mat2, output_quantizer, strict_quantization)

    def int8_matmul(input=None, other=None, output_quantizer=None, strict_quantization=None):
        pa, pb = input.quant_args(), other.quant_args()
        qa, qb = input.raw_data, other.raw_data
        k, n = qb.shape[-2], qb.shape[-1]
        lead = tuple(qa.shape[:-2])
        m = qa.shape[-2]
        qa3 = qa.reshape(-1, m, k).contiguous()
        qbt = qb.reshape(-1, k, n).transpose(1, 2).contiguous()       # [B, N, K]: a no-op for a K-contiguous `other`
        nb = qa3.shape[0]
        out_dtype = pa.dequantize_dtype or torch.float32
        dev = qa.device
        y = torch.empty((nb, m, n), dtype=out_dtype, device=dev)
        sa = pa.scale.detach().reshape(-1)
        oa = None if pa.offset is None else pa.offset.detach().reshape(-1)
        sb = pb.scale.detach().reshape(-1).expand(n).contiguous()      # "per output column" parameters of the GEMM
        ob = None if pb.offset is None else pb.offset.detach().reshape(-1).expand(n).contiguous()
        with C.device_of(dev):
            stream = C.current_stream(dev)
            rs_b = torch.empty((nb, n), dtype=torch.int32, device=dev)
            C.check(C.lib.ffq_rowsum_i8(qbt.data_ptr(), rs_b.data_ptr(), nb * n, k, stream))
            rs_a = None
            if ob is not None:
                rs_a = torch.empty((nb, m), dtype=torch.int32, device=dev)
                C.check(C.lib.ffq_rowsum_i8(qa3.data_ptr(), rs_a.data_ptr(), nb * m, k, stream))
            for i in range(nb):
                C.check(C.lib.ffq_qlinear_w8a8(
                    qa3[i].data_ptr(), qbt[i].data_ptr(), y[i].data_ptr(), C.dtype_tag(out_dtype), m, n, k,
                    sa.data_ptr(), C.ptr(oa), sb.data_ptr(), C.ptr(ob), rs_b[i].data_ptr(),
                    None if rs_a is None else rs_a[i].data_ptr(), None, C.DT_NONE, None, stream))
        _stats["calls_matmul"] = _stats.get("calls_matmul", 0) + nb
        if keepalive is not None:
            keepalive.append((qa3, qbt, y, sa, oa, sb, ob, rs_a, rs_b))
        y = y.reshape(*lead, m, n)
        if output_quantizer is not None:
            y = output_quantizer(y)
        return y

    def int8_mm(input=None, mat2=None, output_quantizer=None, strict_quantization=None):
        return int8_matmul(input, mat2, output_quantizer, strict_quantization)

    return types.SimpleNamespace(
        accepts_w8a8=accepts_w8a8, w8a8_linear=w8a8_linear, accepts_w4a16=accepts_w4a16, w4a16_linear=w4a16_linear,
        accepts_matmul=accepts_matmul, accepts_mm=accepts_mm, accepts_bmm=accepts_bmm, int8_matmul=int8_matmul,
        int8_mm=int8_mm, w4_group=w4_group, params=params)


def register_all(kernels, register, Predicate) -> list:
    """Register one host's kernels with that host's dispatcher; returns the registration hooks."""
    # the dispatcher tries the newest registration first: W8A8 goes last so that an int8 x int8 call is decided by
    # ONE predicate (accepts_w4a16 would evaluate accepts_w8a8 a second time before declining)
    return [
        register("linear", Predicate(kernels.accepts_w4a16), kernels.w4a16_linear),
        register("linear", Predicate(kernels.accepts_w8a8), kernels.w8a8_linear),
        register("matmul", Predicate(kernels.accepts_matmul), kernels.int8_matmul),
        register("mm", Predicate(kernels.accepts_mm), kernels.int8_mm),
        register("bmm", Predicate(kernels.accepts_bmm), kernels.int8_mm),
    ]


_own = None
_hooks: list = []


def own():
    """The kernels bound to this package's classes (built on first use)."""
    global _own
    if _own is None:
        _own = build(_own_host())
    return _own


# module-level names kept for callers and tests
def _accepts(*a, **k):
    return own().accepts_w8a8(*a, **k)


def _accepts_w4a16(*a, **k):
    return own().accepts_w4a16(*a, **k)


def w8a8_linear(*a, **k):
    return own().w8a8_linear(*a, **k)


def w4a16_linear(*a, **k):
    return own().w4a16_linear(*a, **k)


def install() -> None:
    """Register the kernels with ``fastforward_b200.dispatcher`` (idempotent)."""
    global _hooks
    if not _hooks:
        from ..dispatcher import Predicate, register

        _hooks = register_all(own(), register, Predicate)


def uninstall() -> None:
    global _hooks
    for h in _hooks:
        h.remove()
    _hooks = []


def stats() -> Dict[str, Any]:
    return dict(_stats)
