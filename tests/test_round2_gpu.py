"""GPU parity tests for the round-2 pieces: estimator dedupe / memoisation, fused output requantisation, overflow-safe
offset terms, int8 matmul / bmm, freeze_parameters, the cluster (B-multicast) GEMM variants and the BASELINE shapes at
full size.  Everything goes through the public API / the C ABI; the checker is the oracle (oracle/ref_ops.py) or an
exact integer identity."""
import os

import pytest
import torch

import fastforward_b200 as ff
from fastforward_b200 import _cabi as C
from fastforward_b200 import ops
from fastforward_b200.nn import qlinear
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
DEV = "cuda"


# ------------------------------------------------------------------------------------------------------------
# estimator: dedupe of shared inputs, memoisation of unchanged parameters
# ------------------------------------------------------------------------------------------------------------
def _tiny(seed=3):
    import bench_workloads as bw

    model = bw.DecoderStack(bw.TINY, dtype=torch.bfloat16, device=DEV)
    bw.init_weights_(model, seed=seed)
    bw.quantize_for_w8a8(ff, model)
    model.to(DEV)                   # quantizers are created on the CPU (LinearQuantizer's default device)
    qlinear.install()
    return model


def _run(model, batches, **kw):
    est = ff.range_setting.running_minmax(**kw)
    outs = []
    with torch.no_grad(), ff.estimate_ranges(model, est):
        for b in batches:
            outs.append(model(b).clone())
    params = {n: (q.scale.detach().clone(), None if q.offset is None else q.offset.detach().clone())
              for n, q in ff.nn.named_quantizers(model)}
    return outs, params, est.last_stats


def test_dedupe_and_memoisation_are_bit_identical():
    g = torch.Generator().manual_seed(5)
    batches = [torch.randint(0, 1024, (1, 64), generator=g).to(DEV) for _ in range(4)]
    base_out, base_par, base_stats = _run(_tiny(), batches, dedupe=False, memoize_parameters=False)
    assert base_stats["deduped"] == 0 and base_stats["memoized"] == 0 and base_stats["fused"] > 0
    for kw in (dict(dedupe=True, memoize_parameters=False), dict(dedupe=False, memoize_parameters=True),
               dict(dedupe=True, memoize_parameters=True)):
        out, par, stats = _run(_tiny(), batches, **kw)
        if kw["dedupe"]:
            # per layer q/k/v share one launch and gate/up another: 3 of 7 activation steps per layer per batch
            assert stats["deduped"] == 3 * 2 * len(batches), stats
        if kw["memoize_parameters"]:
            assert stats["memoized"] == 7 * 2 * (len(batches) - 1), stats
        for a, b in zip(out, base_out):
            assert torch.equal(a, b), kw
        assert par.keys() == base_par.keys()
        for name in par:
            assert torch.equal(par[name][0], base_par[name][0]), (kw, name)
            assert (par[name][1] is None) == (base_par[name][1] is None)
            if par[name][1] is not None:
                assert torch.equal(par[name][1], base_par[name][1]), (kw, name)


@pytest.mark.parametrize("ahead", [1, 3, 64])
def test_overlapped_parameter_steps_are_bit_identical(ahead):
    """Weights calibrated `ahead` quantizers early on the side stream: outputs and parameters equal the in-line run."""
    g = torch.Generator().manual_seed(6)
    batches = [torch.randint(0, 1024, (1, 64), generator=g).to(DEV) for _ in range(4)]
    base_out, base_par, _ = _run(_tiny(), batches, dedupe=True, memoize_parameters=False)
    out, par, stats = _run(_tiny(), batches, dedupe=True, memoize_parameters=False, overlap_parameters=ahead)
    # 7 weights x 2 layers; the first step records the order, afterwards all but the first weight of a step run ahead
    assert stats["overlapped"] == (7 * 2 - 1) * (len(batches) - 1), stats
    for a, b in zip(out, base_out):
        assert torch.equal(a, b)
    for name in par:
        assert torch.equal(par[name][0], base_par[name][0]), name
        if par[name][1] is not None:
            assert torch.equal(par[name][1], base_par[name][1]), name


def test_overlapped_parameter_steps_under_graph_capture():
    """The side stream joins the capture as a parallel branch; replays equal eager steps on the same inputs."""
    g = torch.Generator().manual_seed(7)
    batches = [torch.randint(0, 1024, (1, 64), generator=g).to(DEV) for _ in range(5)]
    base_out, base_par, _ = _run(_tiny(), batches, memoize_parameters=False)
    model = _tiny()
    est = ff.range_setting.running_minmax(memoize_parameters=False, overlap_parameters=4)
    static = torch.empty_like(batches[0])
    outs = []
    with torch.no_grad(), ff.estimate_ranges(model, est):
        for b in batches[:2]:
            static.copy_(b)
            outs.append(model(static).clone())
        torch.cuda.synchronize()
        cg = torch.cuda.CUDAGraph()
        static.copy_(batches[2])
        with torch.cuda.graph(cg):
            y = model(static)
        # capture does not execute: replay for batch 2, then the remaining batches
        for b in batches[2:]:
            static.copy_(b)
            cg.replay()
            outs.append(y.clone())
    assert est.last_stats["overlapped"] > 0
    for a, b in zip(outs, base_out):
        assert torch.equal(a, b)
    for name, q in ff.nn.named_quantizers(model):
        assert torch.equal(q.scale.detach(), base_par[name][0]), name


def test_overlapped_step_is_dropped_when_the_parameter_changes():
    torch.manual_seed(1)
    lins = torch.nn.Sequential(*[torch.nn.Linear(256, 256, dtype=torch.bfloat16, device=DEV) for _ in range(3)])
    ff.quantize_model(lins)
    for lin in lins:
        lin.weight_quantizer = ff.nn.LinearQuantizer(8, granularity=ff.PerChannel(0), quantized_dtype=torch.int8).to(DEV)
        lin.input_quantizer = ff.nn.LinearQuantizer(8, symmetric=False, quantized_dtype=torch.int8).to(DEV)
    qlinear.install()
    x = torch.randn(4, 16, 256, device=DEV, dtype=torch.bfloat16)
    est = ff.range_setting.running_minmax(memoize_parameters=False, overlap_parameters=2)
    w0 = lins[2].weight.detach().clone()
    hook = lins[1].register_forward_hook(lambda *a: lins[2].weight.mul_(1.5) if hook.live else None)
    hook.live = False
    with torch.no_grad(), ff.strict_quantization(False), ff.estimate_ranges(lins, est):
        lins(x); lins(x)
        hook.live = True                  # the last weight changes AFTER its step was launched ahead
        y = lins(x)
        hook.live = False
    hook.remove()
    w1 = lins[2].weight.detach()
    mn = torch.minimum(w0.min(1).values, w1.min(1).values).cpu()
    mx = torch.maximum(w0.max(1).values, w1.max(1).values).cpu()
    ws, _ = R.parameters_for_range(mn, mx, 8, True, True)
    assert torch.equal(lins[2].weight_quantizer.scale.detach().cpu(), ws)
    # the output used codes of the CHANGED weight: equal to an in-line estimator continuing from the same state
    want = R.quantize_by_tile(w1.cpu(), ws, (1, 256), 8, torch.int8, None)
    with torch.no_grad():
        got = lins[2].weight_quantizer(lins[2].weight).raw_data.cpu()
    assert torch.equal(got, want)
    assert torch.isfinite(y).all()


def test_aliased_quantizers_own_their_parameters_after_the_block():
    model = _tiny()
    batches = [torch.randint(0, 1024, (1, 64)).to(DEV) for _ in range(2)]
    _run(model, batches)
    att = model.layers[0].self_attn
    q, k = att.q_proj.input_quantizer, att.k_proj.input_quantizer
    assert torch.equal(q.scale, k.scale) and q.scale.data_ptr() != k.scale.data_ptr()
    assert q.offset.data_ptr() != k.offset.data_ptr()


def test_memo_is_dropped_when_the_parameter_changes():
    lin = torch.nn.Linear(256, 64, dtype=torch.bfloat16, device=DEV)
    ff.quantize_model(lin)
    lin.weight_quantizer = ff.nn.LinearQuantizer(8, granularity=ff.PerChannel(0), quantized_dtype=torch.int8).to(DEV)
    lin.input_quantizer = ff.nn.LinearQuantizer(8, symmetric=False, quantized_dtype=torch.int8).to(DEV)
    qlinear.install()
    x = torch.randn(4, 16, 256, device=DEV, dtype=torch.bfloat16)
    est = ff.range_setting.running_minmax()
    w0 = lin.weight.detach().clone()
    with torch.no_grad(), ff.strict_quantization(False), ff.estimate_ranges(lin, est):
        lin(x); lin(x)
        lin.weight.mul_(3.0)                     # bumps the version counter: the memo must not be used
        lin(x)
    assert est.last_stats["memoized"] == 1
    w1 = lin.weight.detach()
    mn = torch.minimum(w0.min(1).values, w1.min(1).values).cpu()
    mx = torch.maximum(w0.max(1).values, w1.max(1).values).cpu()
    ws, _ = R.parameters_for_range(mn, mx, 8, True, True)
    assert torch.equal(lin.weight_quantizer.scale.detach().cpu(), ws)


@pytest.mark.parametrize("shape,dtype", [((512, 6), torch.float32), ((3, 5, 1001), torch.float32), ((8, 24, 50), torch.bfloat16)])
def test_fused_step_accepts_rows_that_are_not_whole_vectors(shape, dtype):
    """Per-tensor int8 calibration on shapes whose last dimension is not a multiple of the 16-byte vector: no row
    sums ride along, nothing raises, results equal the oracle's (minmax.py:167)."""
    torch.manual_seed(0)
    x = torch.randn(shape, dtype=dtype, device=DEV)
    if x.numel() % (16 // x.element_size()):
        pytest.skip("numel itself is not a whole number of vectors: the separate kernels handle it")
    q = ff.nn.LinearQuantizer(8, symmetric=False, quantized_dtype=torch.int8).to(DEV)
    with torch.no_grad(), ff.estimate_ranges(q, ff.range_setting.running_minmax):
        out = q(x)
    xc = x.cpu()
    s, o = R.parameters_for_range(xc.min().reshape(1).float(), xc.max().reshape(1).float(), 8, False, True)
    want = R.quantize_by_tile(xc, s, tuple(shape), 8, torch.int8, o)
    assert torch.equal(out.raw_data.cpu(), want)
    assert getattr(out, "_ffq_rowsum", None) is None or shape[-1] % (16 // x.element_size()) == 0


# ------------------------------------------------------------------------------------------------------------
# W8A8 linear: fused output requantisation, wide offsets, cluster variants, full BASELINE shapes
# ------------------------------------------------------------------------------------------------------------
def _gemm(qx, qw, sx, ox, sw, ow, bias=None, out_dtype=torch.bfloat16, requant=None):
    m, k = qx.shape
    n = qw.shape[0]
    st = C.current_stream(qx.device)
    rs_w = torch.empty(n, dtype=torch.int32, device=qx.device)
    C.check(C.lib.ffq_rowsum_i8(qw.data_ptr(), rs_w.data_ptr(), n, k, st))
    rs_x = None
    if ow is not None:
        rs_x = torch.empty(m, dtype=torch.int32, device=qx.device)
        C.check(C.lib.ffq_rowsum_i8(qx.data_ptr(), rs_x.data_ptr(), m, k, st))
    y = torch.empty(m, n, dtype=out_dtype, device=qx.device)
    import ctypes
    C.check(C.lib.ffq_qlinear_w8a8(qx.data_ptr(), qw.data_ptr(), y.data_ptr(), C.dtype_tag(out_dtype), m, n, k,
                                   sx.data_ptr(), C.ptr(ox), sw.data_ptr(), C.ptr(ow), rs_w.data_ptr(), C.ptr(rs_x),
                                   C.ptr(bias), C.dtype_tag(bias.dtype if bias is not None else None),
                                   ctypes.byref(requant) if requant is not None else None, st))
    return y


def _ref_f64(qx, qw, sx, ox, sw, ow, bias=None):
    ok = qx.shape[0] > 16 and qx.shape[1] % 8 == 0 and qw.shape[0] % 8 == 0            # torch._int_mm's shape rules
    acc = torch._int_mm(qx, qw.t()).double() if ok else (qx.double() @ qw.double().t())
    oxr = 0.0 if ox is None else torch.round(ox.double())
    owr = torch.zeros_like(sw, dtype=torch.float64) if ow is None else torch.round(ow.double())
    k = qx.shape[1]
    t = acc + oxr * qw.double().sum(1)[None, :] + owr[None, :] * qx.double().sum(1)[:, None] + k * oxr * owr[None, :]
    y = (sx.double() * sw.double())[None, :] * t
    return y if bias is None else y + bias.double()[None, :]


@pytest.mark.parametrize("cluster", ["", "2", "4", "8"])
@pytest.mark.parametrize("m,n,k", [(2048, 1024, 512), (1024, 768, 4096), (520, 300, 256)])
def test_w8a8_cluster_variants_exact(cluster, m, n, k, monkeypatch):
    """Every cluster size of the pair kernel (B multicast across 1, 2 or 4 pairs) produces the same bits; fp32 output
    of exactly representable products checks the int32 accumulators themselves."""
    if cluster:
        monkeypatch.setenv("FFQ_GEMM_CLUSTER", cluster)
    else:
        monkeypatch.delenv("FFQ_GEMM_CLUSTER", raising=False)
    torch.manual_seed(1)
    qx = torch.randint(-128, 128, (m, k), dtype=torch.int8, device=DEV)
    qw = torch.randint(-128, 128, (n, k), dtype=torch.int8, device=DEV)
    one = torch.ones(1, device=DEV)
    y = _gemm(qx, qw, one, None, torch.ones(n, device=DEV), None, out_dtype=torch.float32)
    want = (qx.double() @ qw.double().t())
    assert torch.equal(y.double(), want)        # |acc| <= 512*2^14 < 2^24: exactly representable in fp32
    # offsets + bias in bf16
    sx = torch.tensor([0.013], device=DEV); ox = torch.tensor([7.4], device=DEV)
    sw = torch.rand(n, device=DEV) * 0.01 + 1e-3; ow = torch.randint(-5, 6, (n,), device=DEV).float()
    b = torch.randn(n, device=DEV)
    y = _gemm(qx, qw, sx, ox, sw, ow, b)
    ref = _ref_f64(qx, qw, sx, ox, sw, ow, b)
    assert torch.allclose(y.double(), ref, rtol=2 ** -7, atol=1e-3)
    assert torch.equal(y, ref.to(torch.float32).to(torch.bfloat16)) or \
        (y.double() - ref).abs().max() <= (ref.abs().max() * 2 ** -7)


def test_w8a8_fused_requant_equals_quantize_of_the_output():
    import ctypes
    torch.manual_seed(2)
    m, n, k = 1536, 640, 1024
    qx = torch.randint(-128, 128, (m, k), dtype=torch.int8, device=DEV)
    qw = torch.randint(-128, 128, (n, k), dtype=torch.int8, device=DEV)
    sx = torch.tensor([0.02], device=DEV); ox = torch.tensor([-3.0], device=DEV)
    sw = torch.rand(n, device=DEV) * 0.01 + 1e-3
    for out_dtype in (torch.bfloat16, torch.float32):
        y = _gemm(qx, qw, sx, ox, sw, None, out_dtype=out_dtype)
        for bits, sym in ((8, False), (4, True)):
            lo, hi = y.float().min().reshape(1) * 0.8, y.float().max().reshape(1) * 0.8      # some clipping
            s = torch.empty(1, device=DEV); o = torch.empty(1, device=DEV)
            ops.parameters_for_range_(lo, hi, bits, sym, True, s, o)
            want = ops.quantize_by_tile(y, s, tuple(y.shape), float(bits), torch.int8, o)
            codes = torch.empty(m, n, dtype=torch.int8, device=DEV)
            rs = torch.zeros(m, dtype=torch.int32, device=DEV)
            rq = C.Requant(s.data_ptr(), o.data_ptr(), float(bits), codes.data_ptr(), rs.data_ptr())
            y2 = _gemm(qx, qw, sx, ox, sw, None, out_dtype=out_dtype, requant=rq)
            assert torch.equal(y2, y)
            assert torch.equal(codes, want), (out_dtype, bits)
            assert torch.equal(rs, want.int().sum(1).to(torch.int32))


def test_qlinear_fuses_the_output_quantizer_through_the_api():
    torch.manual_seed(3)
    lin = torch.nn.Linear(512, 384, dtype=torch.bfloat16, device=DEV)
    ff.quantize_model(lin)
    lin.input_quantizer = ff.nn.LinearQuantizer(8, symmetric=False, quantized_dtype=torch.int8).to(DEV)
    lin.weight_quantizer = ff.nn.LinearQuantizer(8, granularity=ff.PerChannel(0), quantized_dtype=torch.int8).to(DEV)
    lin.output_quantizer = ff.nn.LinearQuantizer(8, symmetric=False, quantized_dtype=torch.int8).to(DEV)
    qlinear.install()
    x = torch.randn(8, 64, 512, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad(), ff.estimate_ranges(lin, ff.range_setting.running_minmax):
        lin(x)                                        # calibrates all three quantizers (output: separate kernels)
    before = qlinear.stats().get("calls_requant_fused", 0)
    with torch.no_grad():
        out = lin(x)                                  # strict quantization is on: every quantizer is live
    assert qlinear.stats().get("calls_requant_fused", 0) == before + 1
    assert isinstance(out, ff.QuantizedTensor) and out.raw_data.dtype == torch.int8
    # the same call without the fusion (an override in front of the output quantizer disables it)
    handle = lin.output_quantizer.register_override(lambda q, cb, a, k: cb(*a, **k))
    with torch.no_grad():
        ref = lin(x)
    handle.remove()
    assert torch.equal(out.raw_data, ref.raw_data)
    assert torch.equal(out.dequantize(), ref.dequantize())
    assert torch.equal(out._ffq_rowsum, out.raw_data.reshape(-1, 384).int().sum(1).to(torch.int32))


def test_w8a8_offsets_far_from_zero_do_not_overflow():
    """Asymmetric ranges far from zero relative to their width give offsets in the tens of thousands; K*ox*ow leaves
    int32 and the kernel adds the corrections in float instead (ffq_qlinear.cu:70)."""
    torch.manual_seed(4)
    m, n, k = 256, 192, 4096
    x = 100.0 + torch.rand(m, k, device=DEV)
    w = 50.0 + torch.rand(n, k, device=DEV)
    sx = torch.empty(1, device=DEV); ox = torch.empty(1, device=DEV)
    ops.parameters_for_range_(x.min().reshape(1), x.max().reshape(1), 8, False, True, sx, ox)
    sw = torch.empty(n, device=DEV); ow = torch.empty(n, device=DEV)
    ops.parameters_for_range_(w.min(1).values, w.max(1).values, 8, False, True, sw, ow)
    assert float(ox.abs()) > 2e4 and float(ow.abs().min()) > 1e4
    qx = ops.quantize_by_tile(x, sx, (m, k), 8.0, torch.int8, ox)
    qw = ops.quantize_by_tile(w, sw, (1, k), 8.0, torch.int8, ow)
    y = _gemm(qx, qw, sx, ox, sw, ow, out_dtype=torch.float32)
    ref = _ref_f64(qx, qw, sx, ox, sw, ow)
    fallback = (ops.dequantize_by_tile(qx, sx, (m, k), ox, torch.float32) @
                ops.dequantize_by_tile(qw, sw, (1, k), ow, torch.float32).t())
    err, err_fb = (y.double() - ref).abs().max(), (fallback.double() - ref).abs().max()
    assert torch.isfinite(y).all() and err <= 4 * err_fb + 1e-3 * ref.abs().max(), (float(err), float(err_fb))


@pytest.mark.parametrize("m,n,k", [(8192, 14336, 4096), (2048, 1024, 8192), (2048, 512, 28672)])
def test_w8a8_full_baseline_shapes_exact(m, n, k):
    """BASELINE configs[3] (8192 x 14336 x 4096) and the Llama-3-70B K sizes: the int32 accumulators are compared
    EXACTLY with the library's int8 GEMM, the epilogue with fp64."""
    torch.manual_seed(5)
    qx = torch.randint(-128, 128, (m, k), dtype=torch.int8, device=DEV)
    qw = torch.randint(-128, 128, (n, k), dtype=torch.int8, device=DEV)
    acc = torch._int_mm(qx, qw.t())
    # scale 2^-8 and fp32 output: y * 256 is the accumulator itself whenever |acc| < 2^24; above that compare in fp64
    s = torch.tensor([2.0 ** -8], device=DEV)
    y = _gemm(qx, qw, s, None, torch.ones(n, device=DEV), None, out_dtype=torch.float32)
    assert torch.equal(y, (acc.float() * (2.0 ** -8)))
    del y
    sx = torch.tensor([0.011], device=DEV); ox = torch.tensor([5.0], device=DEV)
    sw = torch.rand(n, device=DEV) * 0.01 + 1e-3
    b = torch.randn(n, device=DEV)
    y = _gemm(qx, qw, sx, ox, sw, None, b, out_dtype=torch.float32)
    t = acc.double() + 5.0 * qw.double().sum(1)[None, :]
    ref = (0.011 * sw.double())[None, :] * 0 + (sx.double() * sw.double())[None, :] * t + b.double()[None, :]
    assert (y.double() - ref).abs().max() <= 2 ** -22 * ref.abs().max() + 1e-6


def test_cfg1_full_shape_bit_exact():
    """BASELINE configs[0] at full size: 4096 x 4096 fp32, 8-bit PerChannel(0): codes, values, dx bit-exact against the
    oracle (symmetric, and asymmetric with clipping); per-row sums within 4 eps sum|terms| of fp64."""
    torch.manual_seed(0)
    x = torch.randn(4096, 4096)
    g = torch.randn(4096, 4096)
    tile = (1, 4096)
    for symmetric, shrink in ((True, 1.0), (False, 0.5)):
        mn, mx = R.tile_minmax(x, tile)
        s, o = R.parameters_for_range(mn * shrink, mx * shrink, 8, symmetric, True)
        xd, gd, sd = x.to(DEV), g.to(DEV), s.to(DEV)
        od = None if o is None else o.to(DEV)
        q = ops.quantize_by_tile(xd, sd, tile, 8.0, torch.float32, od)
        y = ops.dequantize_by_tile(q, sd, tile, od, torch.float32)
        yf = ops.fake_quantize_by_tile(xd, sd, tile, 8.0, None, od)
        dx, ds, do = ops.quantize_by_tile_backward(xd, gd, sd, tile, 8.0, od)
        rq = R.quantize_by_tile(x, s, tile, 8, torch.float32, o)
        ry = R.dequantize_by_tile(rq, s, tile, o, torch.float32)
        rdx, rds, rdo = R.quantize_by_tile_backward_f64(x, g, s, tile, 8, o)
        assert torch.equal(q.cpu(), rq) and torch.equal(y.cpu(), ry) and torch.equal(yf.cpu(), ry)
        assert torch.equal(dx.cpu(), rdx)
        assert torch.allclose(ds.cpu().double(), rds, rtol=1e-5, atol=1e-3)
        if o is not None:
            assert torch.allclose(do.cpu().double(), rdo, rtol=1e-5, atol=1e-3)


# ------------------------------------------------------------------------------------------------------------
# int8 matmul / bmm through the dispatcher
# ------------------------------------------------------------------------------------------------------------
def _qt(x, bits=8, symmetric=False):
    q = ff.nn.LinearQuantizer(bits, symmetric=symmetric, quantized_dtype=torch.int8).to(x.device)
    q.quantization_range = (x.min(), x.max())
    with torch.no_grad():        # inference: with autograd recording, the scale Parameter would route to the fallback
        return q(x)


@torch.no_grad()
def test_int8_matmul_and_bmm():
    from fastforward_b200.nn import functional as F

    qlinear.install()
    torch.manual_seed(6)
    qh = _qt(torch.randn(2, 4, 192, 128, device=DEV, dtype=torch.bfloat16))          # [B, H, S, D]
    kh = _qt(torch.randn(2, 4, 160, 128, device=DEV, dtype=torch.bfloat16))
    before = qlinear.stats().get("calls_matmul", 0)
    with ff.strict_quantization(False):
        scores = F.matmul(qh, kh.transpose(-1, -2))                                   # QK^T: K-contiguous other
    assert qlinear.stats().get("calls_matmul", 0) == before + 8
    want = qh.dequantize().double() @ kh.dequantize().double().transpose(-1, -2)
    assert scores.shape == (2, 4, 192, 160) and scores.dtype == torch.bfloat16
    assert (scores.double() - want).abs().max() <= 2 ** -7 * want.abs().max()
    p = _qt(torch.rand(8, 192, 160, device=DEV, dtype=torch.bfloat16))
    v = _qt(torch.randn(8, 160, 64, device=DEV, dtype=torch.bfloat16))
    with ff.strict_quantization(False):
        out = F.bmm(p, v)                                                              # attn @ V: N-contiguous other
        out2 = torch.bmm(p, v)                                                         # the __torch_function__ route
    want = p.dequantize().double() @ v.dequantize().double()
    assert (out.double() - want).abs().max() <= 2 ** -7 * want.abs().max()
    assert torch.equal(out, out2)
    a, b = _qt(torch.randn(96, 256, device=DEV)), _qt(torch.randn(256, 80, device=DEV), symmetric=True)
    with ff.strict_quantization(False):
        y = F.mm(a, b)
    want = a.dequantize().double() @ b.dequantize().double()
    assert y.dtype == torch.float32 and (y.double() - want).abs().max() <= 1e-5 * want.abs().max() + 1e-6
    with ff.strict_quantization(True), pytest.raises(ff.QuantizationError):
        F.matmul(a, b)                     # strict quantization without an output quantizer: the fallback's error


def test_w4a16_kernel_steps_aside_for_autograd_and_strict_mode():
    torch.manual_seed(7)
    w = torch.randn(128, 256, device=DEV, dtype=torch.bfloat16)
    wq = ff.nn.LinearQuantizer(4, granularity=ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0),
                               quantized_dtype=torch.int8).to(DEV)
    wq.quantization_range = ops.tile_minmax(w, (1, 128))
    with torch.no_grad():
        qw = wq(w)
    x = torch.randn(32, 256, device=DEV, dtype=torch.bfloat16)
    k = qlinear.own()
    xg = x.clone().requires_grad_()
    with torch.no_grad():
        assert k.accepts_w4a16(input=x, weight=qw, bias=None, output_quantizer=None, strict_quantization=False)
        assert not k.accepts_w4a16(input=x, weight=qw, bias=None, output_quantizer=None, strict_quantization=True)
        assert k.accepts_w4a16(input=xg, weight=qw, bias=None, output_quantizer=None, strict_quantization=False)
    # with autograd recording and something that requires grad (the input, or the quantizer's scale Parameter) the
    # kernel steps aside: the fallback's dequantize + F.linear carries the gradients
    assert not k.accepts_w4a16(input=xg, weight=qw, bias=None, output_quantizer=None, strict_quantization=False)
    assert not k.accepts_w4a16(input=x, weight=qw, bias=None, output_quantizer=None, strict_quantization=False)
    # gradients flow through the fallback
    qlinear.install()
    from fastforward_b200.nn import functional as F
    y = F.linear(xg, qw, strict_quantization=False)
    y.sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad).all()


# ------------------------------------------------------------------------------------------------------------
# freeze_parameters (reference: tests/quantization/test_freeze.py)
# ------------------------------------------------------------------------------------------------------------
def _frozen_model():
    model = torch.nn.Sequential(torch.nn.Linear(64, 64), torch.nn.Linear(64, 32)).to(DEV)
    ff.quantize_model(model)
    ff.find_quantizers(model, "**/[quantizer:parameter/weight]").initialize(
        ff.nn.LinearQuantizer, num_bits=3, granularity=ff.PerChannel(0))
    model.to(DEV)
    with torch.no_grad(), ff.estimate_ranges(model, ff.range_setting.running_minmax), ff.strict_quantization(False):
        model(torch.randn(4, 8, 64, device=DEV))
    return model


def test_freeze_parameters_in_place_matches_two_step():
    from fastforward_b200.quantization import freeze

    torch.manual_seed(8)
    model = _frozen_model()
    want = [layer.weight_quantizer(layer.weight).dequantize().clone() for layer in model]
    meta = [layer.weight_quantizer.quant_metadata for layer in model]
    before = freeze.stats()["fused_in_place"]
    with torch.no_grad(), freeze.freeze_parameters(model):
        model(torch.randn(2, 64, device=DEV))
    assert freeze.stats()["fused_in_place"] == before + 2
    for layer, w, m in zip(model, want, meta):
        assert torch.equal(layer.weight, w)
        assert isinstance(layer.weight_quantizer, ff.nn.QuantizerStub) and layer.weight_quantizer.quant_metadata == m


def test_freeze_parameters_keeps_quantizers_and_respects_disabled_ones():
    from fastforward_b200.quantization import freeze

    model = _frozen_model()
    with torch.no_grad(), freeze.freeze_parameters(model, remove_quantizers=False):
        model(torch.randn(2, 64, device=DEV))
    assert all(type(layer.weight_quantizer) is ff.nn.LinearQuantizer for layer in model)
    assert all(not list(layer.weight_quantizer.overrides) for layer in model)      # hooks are removed on exit
    model = _frozen_model()
    orig = [layer.weight.clone() for layer in model]
    with torch.no_grad(), ff.disable_quantization(model), freeze.freeze_parameters(model):
        model(torch.randn(2, 64, device=DEV))
    assert all(torch.equal(layer.weight, w) for layer, w in zip(model, orig))
    assert all(type(layer.weight_quantizer) is ff.nn.LinearQuantizer for layer in model)
    with freeze.freeze_parameters(torch.nn.Sequential(torch.nn.Linear(2, 2))):
        pass


def test_ops_work_on_a_non_current_device():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    x = torch.randn(64, 256, device="cuda:1")
    assert torch.cuda.current_device() == 0
    mn, mx = ops.tile_minmax(x, (1, 256))
    assert torch.equal(mn.cpu(), x.cpu().min(1).values) and torch.cuda.current_device() == 0


# ------------------------------------------------------------------------------------------------------------
# per-group calibration: one-thread-per-tile kernel, batched whole-model launch
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("g", [64, 128])
@pytest.mark.parametrize("bits,symmetric,allow,qdtype", [(4, True, True, None), (4, False, True, torch.int8), (3, True, False, None),
                                                         (8, True, True, torch.int8)])
def test_group_fake_quant_thread_per_tile_matches_oracle(dtype, g, bits, symmetric, allow, qdtype):
    """calibrate + snap of per-group tiles (the one-thread-per-tile kernel) against the oracle's three steps, incl.
    all-positive tiles next to mixed ones (the deferred one-sided decision), NaN and inf, and the int8-code variant."""
    torch.manual_seed(11)
    w = (torch.randn(96, 1024) * 0.05).to(dtype)
    w[3, :256] = w[3, :256].abs()              # tiles that are entirely non-negative
    w[7, 130] = float("nan")
    w[9, 700] = float("inf")
    tile = (1, g)
    mn, mx = R.tile_minmax(w, tile)
    s, o = R.parameters_for_range(mn.float(), mx.float(), bits, symmetric, allow)
    cd = qdtype or dtype
    want = R.dequantize_by_tile(R.quantize_by_tile(w, s, tile, bits, cd, o), s, tile, o, dtype)
    wd = w.to(DEV)
    nt = w.numel() // g
    scale = torch.empty(nt, device=DEV)
    offset = None if (symmetric and not allow) else torch.empty(nt, device=DEV)
    out = ops.calibrate_fake_quantize_(wd, tile, bits, symmetric, allow, scale, offset, qdtype)
    from conftest import bits_equal
    assert bits_equal(scale.cpu(), s)
    if o is not None:
        assert bits_equal(offset.cpu(), o)
    elif offset is not None:
        assert torch.equal(offset.cpu(), torch.zeros(nt))
    assert bits_equal(out.cpu(), want)
    if bits <= 8:
        run_mn = torch.full((nt,), float("inf"), dtype=dtype, device=DEV); run_mx = -run_mn
        sc2 = torch.empty(nt, device=DEV); of2 = torch.empty(nt, device=DEV)
        codes, _ = ops.calibrate_quantize_(run_mn, run_mx, wd, tile, bits, symmetric, True, sc2, of2)
        s2, o2 = R.parameters_for_range(mn.float(), mx.float(), bits, symmetric, True)
        assert bits_equal(sc2.cpu(), s2) and torch.equal(codes.cpu(), R.quantize_by_tile(w, s2, tile, bits, torch.int8, o2))
        assert bits_equal(run_mn.cpu(), mn) and bits_equal(run_mx.cpu(), mx)


def test_whole_model_batched_fake_quant_equals_per_weight():
    import bench_workloads as bw
    from fastforward_b200.quantization.fuse import calibrate_and_fuse_qdq_weights

    def build():
        model = torch.nn.ModuleList(bw.DecoderLayer(bw.TINY, torch.bfloat16, DEV) for _ in range(2))
        bw.init_weights_(model, seed=5)
        with torch.no_grad():
            model[0].mlp.up_proj.weight.abs_()             # an all-positive weight: the global one-sided decision
        ff.quantize_model(model, extra_conversion=ff.surrogate_quantized_modules(model))
        ff.find_quantizers(model, "**/[quantizer:parameter/weight]").initialize(
            ff.nn.LinearQuantizer, num_bits=4, granularity=ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0))
        return model.to(DEV)

    a, b = build(), build()
    l0 = C.launch_count()
    assert calibrate_and_fuse_qdq_weights(a) == 14
    batched_launches = C.launch_count() - l0
    l0 = C.launch_count()
    calibrate_and_fuse_qdq_weights(b, batched=False)
    assert batched_launches == 2 and C.launch_count() - l0 == 28
    for (na, pa), (nb, pb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert na == nb and torch.equal(pa, pb), na
    # second call: the cached descriptor table is reused, so the call is free of host-to-device copies and can be
    # captured in a CUDA graph; replaying it equals calling the per-weight path on the same (already snapped) weights
    calibrate_and_fuse_qdq_weights(b, batched=False)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        calibrate_and_fuse_qdq_weights(a)
    g.replay(); torch.cuda.synchronize()
    for (na, pa), (nb, pb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(pa, pb), na


# ------------------------------------------------------------------------------------------------------------
# LPBQ scale compression (export/_lpbq.py:131-160): kernel vs the reference's recorded outputs and the oracle
# ------------------------------------------------------------------------------------------------------------
def test_lpbq_encode_matches_reference_goldens():
    from conftest import load_golden
    from fastforward_b200.quantization.lpbq import LPBQProcessor
    for c in load_golden("lpbq"):
        proc = LPBQProcessor(compressed_bw=c["compressed_bw"], decompressed_bw=c["decompressed_bw"])
        enc = proc.generate_lpbq_encoding("w", c["scale"].to(DEV), c["data_shape"], c["tile_size"], c["compressed_bw"])
        assert enc == c["encoding"], (c["data_shape"], c["tile_size"], c["compressed_bw"])
        rows, cols = c["data_shape"]
        block = c["tile_size"][1] if c["orientation"] == "rows" else c["tile_size"][0]
        shape2d = (rows, cols // block) if c["orientation"] == "rows" else (rows // block, cols)
        q, f = proc.grouped_dynamic_quantize(c["scale"].reshape(shape2d).to(DEV), c["grouping"], c["compressed_bw"])
        assert torch.equal(q.cpu().long(), c["int_scale"]) and torch.equal(f.cpu().flatten(), c["float_scale"].flatten())


@pytest.mark.parametrize("shape,axis", [((4096, 32), 0), ((4096, 112), 0), ((32, 4096), 1), ((3, 1), 0), ((1, 5), 1),
                                        ((14336, 32), 0)])
@pytest.mark.parametrize("bw", [4, 7])
def test_lpbq_encode_matches_oracle(shape, axis, bw):
    g = torch.Generator().manual_seed(shape[0] * 31 + shape[1] + bw)
    s = torch.rand(shape, generator=g) * torch.logspace(-4, 1, shape[1]).reshape(1, -1) + 1e-8
    if shape[1 - axis] > 1:
        s[0, 0] = 0.0                                       # a zero scale clamps to 1 (an all-zero channel is 0 / 0: undefined)
    iq, fs = ops.lpbq_encode(s.to(DEV), axis, bw)
    q, f = R.lpbq_grouped_dynamic_quantize(s, axis, bw)
    assert iq.dtype == torch.int32 and torch.equal(iq.cpu().long(), q)
    assert bits_equal_f32(fs.cpu(), f.flatten())


def bits_equal_f32(a, b):
    return a.shape == b.shape and torch.equal(a.view(torch.int32), b.contiguous().view(torch.int32))


def test_lpbq_argument_checks():
    from fastforward_b200.quantization.lpbq import LPBQProcessor
    for kw in (dict(compressed_bw=8), dict(compressed_bw=16, decompressed_bw=8), dict(compressed_bw=0), dict(decompressed_bw=0)):
        with pytest.raises(ValueError):
            LPBQProcessor(**{"compressed_bw": 4, "decompressed_bw": 8, **kw})
    proc = LPBQProcessor()
    s = torch.rand(64 * 128, device=DEV)
    for data_shape, tile, bits in (((64, 128), (1, 1), 4), ((64, 128), (4, 128), 4)):
        with pytest.raises(ValueError, match="not suitable for LPBQ"):
            proc.generate_lpbq_encoding("w", s, data_shape, tile, bits)
    with pytest.raises(ValueError, match="not suitable for LPBQ"):
        LPBQProcessor(8, 16).generate_lpbq_encoding("w", s[:64 * 32], (64, 128), (4, 1), 4)
    with pytest.raises(ValueError, match="not suitable for LPBQ"):
        proc.generate_lpbq_encoding("w", s[:64 * 32], (64, 128), (4, 1), 4, is_symmetric=False)
    with pytest.raises(RuntimeError):
        ops.lpbq_encode(torch.rand(4, 4), 0, 4)                 # host tensor: no CPU path
    q = ff.nn.LinearQuantizer(4, granularity=ff.PerBlock(block_dims=1, block_sizes=16, per_channel_dims=0)).to(DEV)
    w = torch.randn(32, 64, device=DEV)
    with torch.no_grad(), ff.estimate_ranges(q, ff.range_setting.running_minmax):
        q(w)
    enc = proc.encode_quantizer("w", q, w.shape)
    assert enc["block_size"] == 16 and len(enc["scale"]) == 32 and len(enc["per_block_int_scale"]) == 32 * 4
    assert all(1 <= v <= 16 for v in enc["per_block_int_scale"]) and enc["offset"] == [-128.0] * 32


# ------------------------------------------------------------------------------------------------------------
# on-disk formats with parameters that live on the device (tests/test_save_load.py covers the formats on the CPU)
# ------------------------------------------------------------------------------------------------------------
def test_save_and_load_a_calibrated_model_from_the_device(tmp_path):
    model = _tiny()
    batches = [torch.randint(0, 1024, (1, 64)).to(DEV) for _ in range(2)]
    _run(model, batches, memoize_parameters=False)
    path = ff.quantization.save_quantized_model(model, tmp_path / "artifact", name_or_path="tiny")
    fresh = _tiny(seed=9)                              # other weights, other (uncalibrated) quantizers
    ff.quantization.stub_weight_quantizers(fresh)      # some slots hold stubs, some initialised quantizers: both get replaced
    ff.quantization.load_quantized_model(fresh, path, expected_name="tiny", overwrite_policy="overwrite")
    fresh.to(DEV)
    want, got = dict(ff.nn.named_quantizers(model)), dict(ff.nn.named_quantizers(fresh))
    assert want.keys() == got.keys() and len(want) == 28
    for name, q in want.items():
        assert torch.equal(got[name].scale.detach().to(DEV), q.scale.detach()), name
        assert (q.offset is None) == (got[name].offset is None)
        if q.offset is not None:
            assert torch.equal(got[name].offset.detach().to(DEV), q.offset.detach()), name
    x = batches[0]
    with torch.no_grad():
        assert torch.equal(model(x), fresh(x))
