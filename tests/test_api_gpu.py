"""The reference-facing API on the GPU: LinearQuantizer / QuantizedTensor / estimate_ranges /
QuantizedLinear driven exactly like the reference's own tests drive them, checked against vectors
recorded from the unmodified reference (tests/golden/, oracle/make_golden.py)."""
import copy
import pickle

import pytest
import torch

from conftest import bits_equal, load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import fastforward_b200 as ff
    from oracle import ref_ops as R

DEV = "cuda"
GRANS = {
    "per_tensor": lambda: ff.PerTensor(),
    "per_channel0": lambda: ff.PerChannel(0),
    "per_channel1": lambda: ff.PerChannel(1),
    "per_block128": lambda: ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0),
    "per_block32": lambda: ff.PerBlock(block_dims=1, block_sizes=32, per_channel_dims=0),
    "per_tile": lambda: ff.PerTile((4, 16)),
    "per_channel_last": lambda: ff.PerChannel(2),
    "per_block": lambda: ff.PerBlock(block_dims=1, block_sizes=16, per_channel_dims=0),
}


def _sum_close(got, terms_rows, dtype_eps):
    exact = terms_rows.double().sum(1)
    bound = terms_rows.double().abs().sum(1)
    err = (got.detach().cpu().double().reshape(-1) - exact).abs()
    return bool((err <= 4 * dtype_eps * bound + 1e-30).all())


QUANTIZER = load_golden("quantizer")


@pytest.mark.parametrize("i", range(len(QUANTIZER)))
def test_linear_quantizer_matches_reference(i):
    c = QUANTIZER[i]
    q = ff.nn.LinearQuantizer(c["num_bits"], symmetric=c["symmetric"], allow_one_sided=c["allow_one_sided"],
                              granularity=GRANS[c["gran"]](), device=DEV)
    q.quantization_range = (c["range_min"].to(DEV), c["range_max"].to(DEV))
    assert bits_equal(q.scale.detach(), c["scale"])
    if c["offset"] is None:
        assert q.offset is None
    else:
        assert bits_equal(q.offset.detach(), c["offset"])
        assert isinstance(q.offset, torch.nn.Parameter) == c["offset_is_param"]
    x = c["x"].to(DEV).requires_grad_(True)
    qt = q(x)
    assert isinstance(qt, ff.QuantizedTensor) and bits_equal(qt.raw_data, c["q"])
    y = qt.dequantize()
    assert bits_equal(y, c["y"]) and y.dtype == c["x"].dtype
    y.backward(c["grad"].to(DEV))
    assert bits_equal(x.grad, c["dx"])
    _, dsc, doff = R.backward_terms(c["x"], c["grad"], c["scale"], c["tile"], c["num_bits"], c["offset"])
    assert _sum_close(q.scale.grad, dsc, torch.finfo(torch.float32).eps)
    if c["doffset"] is not None:
        assert _sum_close(q.offset.grad, doff, torch.finfo(torch.float32).eps)
    lo, hi = q.quantization_range
    assert bits_equal(lo.detach(), c["range_after"][0]) and bits_equal(hi.detach(), c["range_after"][1])
    # fused path: same bits, same gradients
    x2 = c["x"].to(DEV).requires_grad_(True)
    q.scale.grad = None
    y2 = q.fake_quantize(x2)
    assert bits_equal(y2, c["y"])
    y2.backward(c["grad"].to(DEV))
    assert bits_equal(x2.grad, c["dx"])
    with ff.export_mode(True):
        y3 = q(c["x"].to(DEV))
    assert not isinstance(y3, ff.QuantizedTensor) and bits_equal(y3, c["y"])


MINMAX = load_golden("running_minmax")


@pytest.mark.parametrize("i", range(len(MINMAX)))
@pytest.mark.parametrize("eager", [False, True])
def test_estimate_ranges_matches_reference(i, eager):
    c = MINMAX[i]
    q = ff.nn.LinearQuantizer(8, symmetric=c["symmetric"], granularity=GRANS[c["gran"]](), device=DEV)
    outs = []
    with torch.no_grad(), ff.estimate_ranges(q, ff.range_setting.running_minmax,
                                             disable_quantization=c["disable_quantization"], eager_checks=eager):
        for b in c["batches"]:
            outs.append(q(b.to(DEV)))
    assert list(q.overrides) == []                       # cleanup removed the estimator override
    assert bits_equal(q.scale.detach(), c["scale"]) and bits_equal(q.offset.detach(), c["offset"])
    last = outs[-1]
    if c["disable_quantization"]:
        assert not isinstance(last, ff.QuantizedTensor) and bits_equal(last, c["last_raw"])
    else:
        assert bits_equal(last.raw_data, c["last_raw"])
    lo, hi = q.quantization_range
    assert bits_equal(lo.detach(), c["range"][0]) and bits_equal(hi.detach(), c["range"][1])


def test_estimate_ranges_inf_raises():
    q = ff.nn.LinearQuantizer(8, device=DEV)
    with pytest.raises(NotImplementedError, match="Infinite"):      # deferred to the end of the block
        with ff.estimate_ranges(q, ff.range_setting.running_minmax):
            q(torch.tensor([1.0, float("inf")], device=DEV))
    q2 = ff.nn.LinearQuantizer(8, device=DEV)
    with pytest.raises(NotImplementedError, match="Infinite"):      # reference timing: inside the step
        with ff.estimate_ranges(q2, ff.range_setting.running_minmax, eager_checks=True):
            q2(torch.tensor([1.0, float("-inf")], device=DEV))
            raise AssertionError("the step above must raise")


def test_smoothed_minmax():
    torch.manual_seed(0)
    q = ff.nn.LinearQuantizer(8, symmetric=False, granularity=ff.PerChannel(0), device=DEV)
    batches = [torch.randn(8, 32) * (i + 1) for i in range(4)]
    with torch.no_grad(), ff.estimate_ranges(q, ff.range_setting.smoothed_minmax, gamma=0.3):
        for b in batches:
            q(b.to(DEV))
    mn = mx = None
    for b in batches:
        mn, mx = R.smoothed_minmax_step(mn, mx, b, (1, 32), 0.3)
    s, o = R.parameters_for_range(mn, mx, 8, False, True)
    assert torch.allclose(q.scale.detach().cpu(), s, rtol=1e-6) and torch.allclose(q.offset.detach().cpu(), o, rtol=1e-5, atol=1e-4)


LINEAR = load_golden("linear")


@pytest.mark.parametrize("i", range(len(LINEAR)))
def test_quantized_linear_fallback_matches_reference(i):
    c = LINEAR[i]
    dt = c["x"].dtype
    lin = torch.nn.Linear(c["k"], c["n"], bias=c["bias"] is not None, dtype=dt)
    with torch.no_grad():
        lin.weight.copy_(c["w"])
        if c["bias"] is not None:
            lin.bias.copy_(c["bias"])
    ff.quantize_model(lin)
    assert isinstance(lin, ff.nn.QuantizedLinear)
    wsym = c["w_offset"] is None or bool((c["w_offset"] == 0).all())
    lin.input_quantizer = ff.nn.LinearQuantizer(8, symmetric=False, quantized_dtype=torch.int8)
    lin.weight_quantizer = ff.nn.LinearQuantizer(8, symmetric=wsym, granularity=ff.PerChannel(0),
                                                 quantized_dtype=torch.int8)
    lin.to(DEV)
    x = c["x"].to(DEV)
    lin.input_quantizer.quantization_range = (x.min(), x.max())
    lin.weight_quantizer.quantization_range = (lin.weight.min(1).values, lin.weight.max(1).values)
    assert bits_equal(lin.input_quantizer.scale.detach(), c["x_scale"])
    assert bits_equal(lin.weight_quantizer.scale.detach(), c["w_scale"])
    with torch.no_grad(), ff.strict_quantization(False), ff.dispatcher.register("linear", None, ff.nn.functional.fallback.linear):
        y = lin(x)
    tol = dict(rtol=1e-4, atol=1e-4) if dt is torch.float32 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(y.cpu(), c["y"], **tol)
    with pytest.raises(ff.QuantizationError):           # strict mode: an output quantizer stub returns a plain tensor
        # (set explicitly: the flag is process-wide and other test modules switch it off)
        with ff.strict_quantization(True), ff.dispatcher.register("linear", None, ff.nn.functional.fallback.linear):
            ff.nn.functional.linear(x, lin.weight, None, output_quantizer=None)


def test_quantized_tensor_behaviour():
    torch.manual_seed(0)
    x = torch.randn(6, 8, device=DEV)
    q = ff.nn.LinearQuantizer(8, granularity=ff.PerChannel(0), device=DEV)
    q.quantization_range = (x.min(1).values, x.max(1).values)
    qt = q(x)
    assert qt.shape == x.shape and qt.dtype == x.dtype and qt.is_cuda and qt.dim() == 2 and qt.numel() == 48
    deq = qt.dequantize()
    assert isinstance(qt.clone(), ff.QuantizedTensor) and bits_equal(qt.clone().dequantize(), deq)
    assert isinstance(qt.detach(), ff.QuantizedTensor) and qt.contiguous() is qt
    cpu = qt.cpu()
    assert isinstance(cpu, ff.QuantizedTensor) and cpu.quant_args().scale.device.type == "cpu"
    assert bits_equal(cpu.cuda().dequantize(), deq)
    assert bits_equal(qt.float(), deq) and qt.to(torch.float16).dtype == torch.float16
    assert not isinstance(qt.half(), ff.QuantizedTensor)
    dc = copy.deepcopy(qt.detach())
    assert isinstance(dc, ff.QuantizedTensor) and bits_equal(dc.dequantize(), deq)
    rt = pickle.loads(pickle.dumps(qt.detach().cpu()))
    assert isinstance(rt, ff.QuantizedTensor) and torch.equal(rt.raw_data, qt.raw_data.cpu())
    with pytest.raises(ff.QuantizationError):            # strict: no implicit dequantization
        qt + 1
    with ff.strict_quantization(False):
        assert bits_equal(qt + 1, deq + 1)
        assert bits_equal(torch.add(qt, 1), deq + 1)
    with pytest.raises(NotImplementedError):
        qt.add_(1)
    with ff.dispatcher.register("add", None, lambda a, b: "custom"):
        assert (qt + 1) == "custom" and torch.add(qt, 1) == "custom"
    assert qt.view(6, 8) is qt


def test_functional_frontends_and_dynamic():
    torch.manual_seed(1)
    A = ff.quantization.affine
    x = torch.randn(16, 64, device=DEV)
    s = torch.tensor([0.05], device=DEV)
    qt = A.quantize_per_tensor(x, s, None, 4)
    assert bits_equal(qt.raw_data, R.quantize_by_tile(x.cpu(), s.cpu(), x.shape, 4, x.dtype, None))
    sc = torch.rand(16, device=DEV) * 0.1 + 0.01
    oc = torch.randn(16, device=DEV)
    qc = A.quantize_per_channel(x, sc, oc, axis=0, num_bits=8, output_dtype=torch.int8)
    assert bits_equal(qc.raw_data, R.quantize_by_tile(x.cpu(), sc.cpu(), (1, 64), 8, torch.int8, oc.cpu()))
    assert bits_equal(qc.dequantize(), R.dequantize_by_tile(qc.raw_data.cpu(), sc.cpu(), (1, 64), oc.cpu(), torch.float32))
    sb = torch.rand(64, device=DEV) * 0.1 + 0.01
    qb = A.quantize_per_block(x, sb, torch.zeros(64, device=DEV), channel_axis=0, block_axis=1, block_size=16, num_bits=4)
    assert bits_equal(qb.raw_data, R.quantize_by_tile(x.cpu(), sb.cpu(), (1, 16), 4, x.dtype, torch.zeros(64)))
    qd = A.dynamic.quantize_per_channel(x, axis=0, num_bits=8, symmetric=True)
    rq, rs, ro = R.quantize_dynamic_by_tile(x.cpu(), (1, 64), 8.0, True, True, x.dtype)
    assert bits_equal(qd.raw_data, rq) and bits_equal(qd.quant_args().scale, rs)
    # dynamic == static with the exact range (reference tests/quantization/test_dynamic.py:12-30)
    lq = ff.nn.LinearQuantizer(8, granularity=ff.PerChannel(0), device=DEV)
    lq.quantization_range = (x.min(1).values, x.max(1).values)
    assert bits_equal(lq(x).dequantize(), qd.dequantize())
    # python-scalar scale: converted to the data dtype (affine/_autograd.py:35-37)
    assert bits_equal(A.quantize_per_tensor(x, 0.05, None, 4).raw_data, A.quantize_per_tensor(x, torch.tensor(0.05, device=DEV), None, 4).raw_data)


def test_fuse_qdq_weights_matches_two_step_and_is_idempotent():
    """reference tests/quantization/test_fuse.py:51-117: fused weights == quantizer(w).dequantize(), idempotent."""
    from fastforward_b200.quantization.fuse import calibrate_weight_quantizers, fuse_qdq_weights

    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(256, 128, bias=False), torch.nn.Linear(128, 64)).to(torch.bfloat16)
    ff.quantize_model(model)
    ff.find_quantizers(model, "**/[quantizer:parameter/weight]").initialize(
        ff.nn.LinearQuantizer, num_bits=4, granularity=ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0))
    model.to(DEV)
    assert calibrate_weight_quantizers(model) == 2
    expect = []
    for lin in model:
        w = lin.weight.detach().cpu()
        rows = R.tile_rows(w, (1, 128))
        s, o = R.parameters_for_range(rows.min(1).values, rows.max(1).values, 4, True, True)
        o = torch.zeros_like(s) if o is None else o
        assert bits_equal(lin.weight_quantizer.scale.detach(), s)
        expect.append(R.dequantize_by_tile(R.quantize_by_tile(w, s, (1, 128), 4, w.dtype, o), s, (1, 128), o, w.dtype))
    two_step = [lin.weight_quantizer(lin.weight).dequantize() for lin in model]
    assert fuse_qdq_weights(model) == 2
    for lin, want, ts in zip(model, expect, two_step):
        assert bits_equal(lin.weight.detach(), want) and bits_equal(lin.weight.detach(), ts)
    before = [lin.weight.detach().clone() for lin in model]
    fuse_qdq_weights(model, stub_quantizers=True)
    for lin, b in zip(model, before):
        assert bits_equal(lin.weight.detach(), b) and lin.weight_quantizer.is_stub()
    # sharding: rank r of 2 touches only its targets
    m2 = torch.nn.Sequential(*[torch.nn.Linear(128, 128, bias=False) for _ in range(4)])
    ff.quantize_model(m2)
    ff.find_quantizers(m2, "**/[quantizer:parameter/weight]").initialize(ff.nn.LinearQuantizer, num_bits=4, granularity=ff.PerChannel(0))
    m2.to(DEV)
    calibrate_weight_quantizers(m2)
    orig = [lin.weight.detach().clone() for lin in m2]
    assert fuse_qdq_weights(m2, rank=1, world_size=2) == 2
    changed = [not torch.equal(lin.weight.detach(), o) for lin, o in zip(m2, orig)]
    assert changed == [False, True, False, True]


def test_dynamic_linear_quantizer_module():
    """reference tests/nn/test_dynamic_linear_quantizer.py: dynamic == static with the exact per-call range."""
    torch.manual_seed(3)
    x = torch.randn(4, 8, 32, device=DEV, dtype=torch.bfloat16)
    dq = ff.nn.DynamicLinearQuantizer(8, granularity=ff.PerChannel(2), symmetric=False)
    out = dq(x)
    assert isinstance(out, ff.QuantizedTensor)
    rq, rs, ro = R.quantize_dynamic_by_tile(x.cpu(), (4, 8, 1), 8.0, False, True, x.dtype)
    assert bits_equal(out.raw_data, rq) and bits_equal(out.quant_args().scale, rs) and bits_equal(out.quant_args().offset, ro)
    assert bits_equal(out.dequantize(), R.dequantize_by_tile(rq, rs, (4, 8, 1), ro, x.dtype))
    xg = torch.randn(16, 16, device=DEV, requires_grad=True)
    ff.nn.DynamicLinearQuantizer(4)(xg).dequantize().sum().backward()
    assert torch.equal(xg.grad, torch.ones_like(xg))          # identity backward (affine/_autograd.py:125-133)
    with pytest.raises(ff.QuantizationError):
        ff.nn.DynamicLinearQuantizer(8)(torch.empty(0, 4, device=DEV))


def test_continue_calibration_from_existing_fp32_range():
    """minmax.py:198-200: an estimator created on an initialised quantizer continues from its (fp32)
    range; with bf16 data torch.min promotes to fp32 -- no conversion pass over the data here."""
    torch.manual_seed(5)
    q = ff.nn.LinearQuantizer(8, symmetric=False, granularity=ff.PerChannel(0), device=DEV)
    a = torch.randn(8, 64, dtype=torch.bfloat16)
    b = torch.randn(8, 64, dtype=torch.bfloat16) * 2
    with torch.no_grad(), ff.estimate_ranges(q, ff.range_setting.running_minmax):
        q(a.to(DEV))
    lo0, hi0 = (t.detach().cpu() for t in q.quantization_range)
    with torch.no_grad(), ff.estimate_ranges(q, ff.range_setting.running_minmax):
        q(b.to(DEV))
    mn, mx = R.running_minmax_step(lo0, hi0, b, (1, 64))          # fp32 running range, bf16 data
    s, o = R.parameters_for_range(mn, mx, 8, False, True)
    assert bits_equal(q.scale.detach(), s) and bits_equal(q.offset.detach(), o)


MSE_GRID = load_golden("mse_grid")


def _mse_grid_run(c, **estimator_kwargs):
    q = ff.nn.LinearQuantizer(c["num_bits"], symmetric=c["symmetric"], granularity=GRANS[c["gran"]](), device=DEV)
    with torch.no_grad(), ff.estimate_ranges(q, ff.range_setting.mse_grid, num_candidates=c["num_candidates"],
                                             **estimator_kwargs):
        for b in c["batches"]:
            q(b.to(DEV))
        step = next(iter(q.overrides))
    return q, step


def _check_mse_grid(c, q, step):
    f32 = c["batches"][0].dtype is torch.float32
    # the grid is parameter-sized arithmetic on per-tile extrema: bit-exact
    assert bits_equal(step.min_threshold, c["min_threshold"]) and bits_equal(step.max_threshold, c["max_threshold"])
    # accumulated errors: same terms, another summation order (and one bf16 rounding per batch)
    tol = dict(rtol=2e-5, atol=1e-7) if f32 else dict(rtol=2e-2, atol=1e-3)
    got, want = step.cumulative_error.cpu(), c["cumulative_error"]
    torch.testing.assert_close(got, want, **tol)
    assert bool((got[c["num_candidates"]:] == 0).all())    # rows the reference never evaluates stay 0
    # selected parameters: bit-exact wherever the arg-min is decided by more than the tolerance
    best_got, best_want = got.min(0).indices, want.min(0).indices
    same = best_got == best_want
    cols = torch.arange(want.shape[1])
    gap = (want[best_got, cols] - want[best_want, cols]).abs()
    assert bool((gap[~same] <= tol["rtol"] * want[best_want, cols][~same].abs() + tol["atol"]).all())
    assert bits_equal(q.scale.detach().cpu()[same], c["scale"][same])
    assert bits_equal(q.offset.detach().cpu()[same], c["offset"][same])


@pytest.mark.parametrize("i", range(len(MSE_GRID)))
def test_mse_grid_matches_reference(i):
    """range_setting/min_error.py through estimate_ranges, fused kernel where the tiles are runs."""
    c = MSE_GRID[i]
    q, step = _mse_grid_run(c)
    assert step._fused == (c["gran"] != "per_channel_last")
    _check_mse_grid(c, q, step)


@pytest.mark.parametrize("i", [0, 5, 17, 33, 50, 63])
def test_mse_grid_custom_error_fn_takes_candidate_loop(i):
    c = MSE_GRID[i]

    def my_mse(quantized, original):
        return torch.mean((quantized - original) ** 2, dim=1)

    q, step = _mse_grid_run(c, error_fn=my_mse)
    assert step._fused is False
    _check_mse_grid(c, q, step)


def test_mse_grid_large_tiles_and_policy():
    """Tiles longer than one warp segment (two-stage sum) against the oracle; update policy honoured."""
    g = torch.Generator().manual_seed(5)
    for dtype, shape, gran, tile in [(torch.float32, (6, 5000), "per_channel0", (1, 5000)),
                                     (torch.bfloat16, (3, 64, 1000), "per_tensor", (3, 64, 1000)),
                                     (torch.float16, (16, 4096), "per_block128", (1, 128))]:
        x = torch.randn(shape, generator=g).to(dtype)
        for symmetric in (True, False):
            q = ff.nn.LinearQuantizer(4, symmetric=symmetric, granularity=GRANS[gran](), device=DEV)
            calls = []
            ntiles = x.numel() // torch.Size(tile).numel()
            q.quantization_range = (-torch.ones(ntiles, device=DEV), torch.ones(ntiles, device=DEV))
            before = q.scale.detach().clone()
            with torch.no_grad(), ff.estimate_ranges(q, ff.range_setting.mse_grid, num_candidates=16,
                                                     update_range_policy=lambda est, n: calls.append(n) or n == 2):
                q(x.to(DEV))
                assert torch.equal(q.scale.detach(), before)      # policy said "not yet"
                q(x.to(DEV))
                step = next(iter(q.overrides))
            assert calls == [1, 2] and step._fused
            lo, hi = R.uniform_search_grid(x, tile, symmetric, 16)
            assert bits_equal(step.min_threshold, lo) and bits_equal(step.max_threshold, hi)
            want = 2 * R.mse_grid_errors(x, tile, lo, hi, 4, symmetric, True, num_candidates=16).float()
            tol = dict(rtol=1e-4, atol=1e-7) if dtype is torch.float32 else dict(rtol=3e-2, atol=1e-3)
            torch.testing.assert_close(step.cumulative_error.cpu().float(), want, **tol)
