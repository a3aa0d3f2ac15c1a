"""W8A8 tensor-core linear (dispatcher "linear") against the reference's dequantize-then-float
fallback and the float64 yardstick.

Tolerance (stated, as BASELINE.json asks): the kernel accumulates exactly in int32, so its only
error is the final float rounding; we require
    |y_kernel - y_f64| <= |y_fallback - y_f64| + 2 eps*(|y|+|bias|)      elementwise,
i.e. never worse than the reference path, plus rtol 1e-3 (fp32) / 2e-2 (bf16) against the
fallback itself (absolute part: the same factor times max|y|)."""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import fastforward_b200 as ff
    from fastforward_b200.nn import qlinear
    from oracle import ref_ops as R

DEV = "cuda"


def _run_case(m, k, n, dt, w_sym, bias, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(m, k, generator=g).to(dt)
    w = (torch.randn(n, k, generator=g) * 0.05).to(dt)
    b = (torch.randn(n, generator=g) * 0.1).to(dt) if bias else None
    lin = torch.nn.Linear(k, n, bias=bias, dtype=dt)
    with torch.no_grad():
        lin.weight.copy_(w)
        if bias:
            lin.bias.copy_(b)
    ff.quantize_model(lin)
    lin.input_quantizer = ff.nn.LinearQuantizer(8, symmetric=False, quantized_dtype=torch.int8)
    lin.weight_quantizer = ff.nn.LinearQuantizer(8, symmetric=w_sym, granularity=ff.PerChannel(0), quantized_dtype=torch.int8)
    lin.to(DEV)
    xc = x.to(DEV)
    lin.input_quantizer.quantization_range = (xc.min(), xc.max())
    lin.weight_quantizer.quantization_range = (lin.weight.min(1).values, lin.weight.max(1).values)
    with torch.no_grad():
        before = qlinear.stats()["calls"]
        qlinear.install()
        try:
            y = lin(xc)
        finally:
            qlinear.uninstall()
        assert qlinear.stats()["calls"] == before + 1, "the tensor-core kernel was not dispatched"
        y_fb = lin(xc)                                      # reference fallback path (nothing registered)
        xq, wq = lin.input_quantizer(xc), lin.weight_quantizer(lin.weight)
    px, pw = xq.quant_args(), wq.quant_args()
    y64 = R.exact_linear_f64(xq.raw_data.cpu(), px.scale.detach().cpu(), px.offset.detach().cpu(), (m, k),
                             wq.raw_data.cpu(), pw.scale.detach().cpu(),
                             None if pw.offset is None else pw.offset.detach().cpu(), (1, k),
                             None if b is None else b)
    assert y.dtype == dt and y.shape == (m, n)
    yk, yf = y.double().cpu(), y_fb.double().cpu()
    bmag = 0.0 if b is None else b.double().abs()[None, :]
    ulp = torch.finfo(dt).eps * (y64.abs() + bmag).clamp_min(1e-3)
    err_k, err_f = (yk - y64).abs(), (yf - y64).abs()
    assert bool((err_k <= err_f + 2 * ulp).all()), f"kernel worse than fallback: {float((err_k - err_f - 2 * ulp).max())}"
    assert float(err_k.max()) <= float(err_f.max()) + float(ulp.max())
    # against the fallback itself: rtol 1e-3 (fp32) / 2e-2 (bf16: the fallback multiplies bf16-rounded
    # operands), absolute part scaled by the output magnitude
    rtol = 1e-3 if dt is torch.float32 else 2e-2
    torch.testing.assert_close(y.float().cpu(), y_fb.float().cpu(), rtol=rtol, atol=rtol * float(y_fb.abs().max()))


@pytest.mark.parametrize("m,k,n", [(128, 128, 256), (256, 512, 512), (300, 1024, 700), (17, 96, 40), (2048, 4096, 1024),
                                   (129, 4096 + 16, 257),
                                   # 224-column pair tiles (chosen when they even out the waves), incl. a ragged last tile
                                   (2048, 256, 14336), (2048, 128, 14336 - 96)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16] if torch.cuda.is_available() else [])
@pytest.mark.parametrize("w_sym,bias", [(True, False), (False, True)])
def test_w8a8_linear(m, k, n, dt, w_sym, bias):
    _run_case(m, k, n, dt, w_sym, bias)


def test_w8a8_3d_input_and_predicate():
    torch.manual_seed(0)
    lin = torch.nn.Linear(256, 512, bias=False, dtype=torch.bfloat16)
    ff.quantize_model(lin)
    lin.input_quantizer = ff.nn.LinearQuantizer(8, symmetric=False, quantized_dtype=torch.int8)
    lin.weight_quantizer = ff.nn.LinearQuantizer(8, granularity=ff.PerChannel(0), quantized_dtype=torch.int8)
    lin.to(DEV)
    x = torch.randn(2, 64, 256, device=DEV, dtype=torch.bfloat16)
    lin.input_quantizer.quantization_range = (x.min(), x.max())
    lin.weight_quantizer.quantization_range = (lin.weight.min(1).values, lin.weight.max(1).values)
    qlinear.install()
    try:
        with torch.no_grad():
            n0 = qlinear.stats()["calls"]
            y = lin(x)
            assert y.shape == (2, 64, 512) and qlinear.stats()["calls"] == n0 + 1
            # float codes (default quantized_dtype) are outside the predicate: fallback, not the kernel
            lin.input_quantizer.quantized_dtype = None
            y2 = lin(x)
            assert qlinear.stats()["calls"] == n0 + 1
            torch.testing.assert_close(y.float(), y2.float(), rtol=2e-2, atol=2e-2)
    finally:
        qlinear.uninstall()


def test_golden_linear_cases_through_kernel():
    for c in load_golden("linear"):
        px_s, px_o = c["x_scale"].to(DEV), c["x_offset"].to(DEV)
        ctx_x = ff.quantization.affine.quantization_context(px_s, px_o, ff.PerTensor(), 8, torch.int8, c["x"].dtype)
        ctx_w = ff.quantization.affine.quantization_context(
            c["w_scale"].to(DEV), None if c["w_offset"] is None else c["w_offset"].to(DEV), ff.PerChannel(0), 8,
            torch.int8, c["w"].dtype)
        xq, wq = ctx_x.attach(c["x_codes"].to(DEV)), ctx_w.attach(c["w_codes"].to(DEV))
        b = None if c["bias"] is None else c["bias"].to(DEV)
        if c["k"] % 16:
            continue
        y = qlinear.w8a8_linear(input=xq, weight=wq, bias=b, output_quantizer=None, strict_quantization=False)
        tol = 1e-4 if c["x"].dtype is torch.float32 else 2e-2
        torch.testing.assert_close(y.cpu().float(), c["y"].float(), rtol=tol, atol=tol)


# ------------------------------------------------------------------------------------------------
# W4A16: weight-only quantized linear.  The in-kernel dequantisation is dequantize_by_tile's
# arithmetic, so (a) with one-hot activations the output IS the dequantized weight, bit for bit, and
# (b) for real activations the only difference to the reference's fallback (dequantize + F.linear)
# is the fp32 accumulation order: |y_kernel - y_f64| <= |y_fallback - y_f64| + 2 ulp(dtype) elementwise
# and rtol 2e-2 against the fallback itself (tolerance stated as BASELINE.json asks).
# ------------------------------------------------------------------------------------------------
def _w4_layer(k, n, dt, gran, w_sym, bias, bits=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    lin = torch.nn.Linear(k, n, bias=bias, dtype=dt)
    with torch.no_grad():
        lin.weight.copy_((torch.randn(n, k, generator=g) * 0.05).to(dt))
        if bias:
            lin.bias.copy_((torch.randn(n, generator=g) * 0.1).to(dt))
    ff.quantize_model(lin)
    lin.weight_quantizer = ff.nn.LinearQuantizer(bits, symmetric=w_sym, granularity=gran, quantized_dtype=torch.int8)
    lin.to(DEV)
    with torch.no_grad(), ff.estimate_ranges(lin.weight_quantizer, ff.range_setting.running_minmax):
        lin.weight_quantizer(lin.weight)
    return lin


def _w4_grans(k):
    return {
        "g128": ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0),
        "g64": ff.PerBlock(block_dims=1, block_sizes=64, per_channel_dims=0),
        "pc": ff.PerChannel(0),
        "pt": ff.PerTensor(),
    }


def _dispatch_w4(lin, x):
    # weight-only quantization: the reference's fallback refuses float inputs under strict mode
    with torch.no_grad(), ff.strict_quantization(False):
        before = qlinear.stats().get("calls_w4a16", 0)
        qlinear.install()
        limit, qlinear.W4A16_MAX_ROWS = qlinear.W4A16_MAX_ROWS, None     # exercise the kernel at every size
        try:
            y = lin(x)
        finally:
            qlinear.W4A16_MAX_ROWS = limit
            qlinear.uninstall()
        assert qlinear.stats().get("calls_w4a16", 0) == before + 1, "the W4A16 kernel was not dispatched"
        return y, lin(x)            # (kernel, reference fallback: nothing registered)


@pytest.mark.parametrize("k,n,gran", [(128, 256, "g128"), (512, 300, "g64"), (1024, 40, "pc"), (256, 513, "pt"), (4096, 256, "g128")])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16] if torch.cuda.is_available() else [])
@pytest.mark.parametrize("w_sym", [True, False])
def test_w4a16_one_hot_activations_give_the_dequantized_weight_bit_exact(k, n, gran, dt, w_sym):
    lin = _w4_layer(k, n, dt, _w4_grans(k)[gran], w_sym, bias=False)
    x = torch.eye(k, dtype=dt, device=DEV)
    y, y_fb = _dispatch_w4(lin, x)
    with torch.no_grad():
        w_deq = lin.weight_quantizer(lin.weight).dequantize()
    assert y.dtype == dt and torch.equal(y, w_deq.t()) and torch.equal(y_fb, y)


@pytest.mark.parametrize("m,k,n,gran", [(256, 512, 512, "g128"), (300, 1024, 700, "g64"), (17, 128, 40, "pc"),
                                        (2048, 4096, 1024, "g128"), (129, 4096 + 64, 257, "pc"), (64, 256, 256, "pt"),
                                        # two 256-row tiles per dequantized weight stage (M > 256), ragged in M and N
                                        (513, 256, 512, "g128"), (1000, 512, 300, "g64"), (257, 128, 264, "pc")])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16] if torch.cuda.is_available() else [])
@pytest.mark.parametrize("w_sym,bias", [(True, False), (False, True)])
def test_w4a16_linear(m, k, n, gran, dt, w_sym, bias):
    lin = _w4_layer(k, n, dt, _w4_grans(k)[gran], w_sym, bias, seed=m + n)
    x = torch.randn(m, k, generator=torch.Generator().manual_seed(7)).to(dt).to(DEV)
    y, y_fb = _dispatch_w4(lin, x)
    with torch.no_grad():
        w_deq = lin.weight_quantizer(lin.weight).dequantize()
    y64 = torch.nn.functional.linear(x.double(), w_deq.double(), None if lin.bias is None else lin.bias.double()).cpu()
    assert y.dtype == dt and y.shape == (m, n)
    yk, yf = y.double().cpu(), y_fb.double().cpu()
    ulp = torch.finfo(dt).eps * y64.abs().clamp_min(1e-2)
    err_k, err_f = (yk - y64).abs(), (yf - y64).abs()
    assert bool((err_k <= err_f + 2 * ulp).all()), f"kernel worse than fallback by {float((err_k - err_f - 2 * ulp).max())}"
    torch.testing.assert_close(y.float().cpu(), y_fb.float().cpu(), rtol=2e-2, atol=2e-2 * float(y_fb.abs().max()))


def test_w4a16_quantized_16bit_activations_3d_input_and_predicate():
    """cfg5 recipe: A16 per-tensor asymmetric (fp32 codes, bf16 data) x W4 g=128; 3-D input; the
    predicate leaves everything it does not implement to the fallback."""
    dt = torch.bfloat16
    lin = _w4_layer(512, 384, dt, _w4_grans(512)["g128"], False, True)
    lin.input_quantizer = ff.nn.LinearQuantizer(16, symmetric=False, quantized_dtype=torch.float32).to(DEV)
    x = torch.randn(2, 50, 512, generator=torch.Generator().manual_seed(3)).to(dt).to(DEV)
    lin.input_quantizer.quantization_range = (x.min(), x.max())
    y, y_fb = _dispatch_w4(lin, x)
    assert y.shape == (2, 50, 384)
    torch.testing.assert_close(y.float(), y_fb.float(), rtol=2e-2, atol=2e-2 * float(y_fb.abs().max()))
    with torch.no_grad():
        wq = lin.weight_quantizer(lin.weight)
        assert qlinear._accepts_w4a16(input=x, weight=wq, bias=None)
        assert not qlinear._accepts_w4a16(input=x.float(), weight=wq, bias=None)            # fp32 activations
        assert qlinear._accepts_w4a16(input=x.repeat(8, 1, 1), weight=wq, bias=None)        # 800 rows: no row limit any more
        assert not qlinear._accepts_w4a16(input=x[..., :448], weight=wq, bias=None)         # K mismatch
        odd = ff.nn.LinearQuantizer(4, granularity=ff.PerBlock(block_dims=1, block_sizes=32, per_channel_dims=0),
                                    quantized_dtype=torch.int8).to(DEV)
        odd.quantization_range = (-torch.ones(384 * 16, device=DEV), torch.ones(384 * 16, device=DEV))
        assert not qlinear._accepts_w4a16(input=x, weight=odd(lin.weight), bias=None)       # group of 32 < one k-block
        pcl = ff.nn.LinearQuantizer(4, granularity=ff.PerChannel(1), quantized_dtype=torch.int8).to(DEV)
        pcl.quantization_range = (-torch.ones(512, device=DEV), torch.ones(512, device=DEV))
        assert not qlinear._accepts_w4a16(input=x, weight=pcl(lin.weight), bias=None)       # per input channel
