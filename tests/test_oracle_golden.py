"""Pin the oracle (oracle/ref_ops.py) to the reference: every vector recorded from the
unmodified reference (tests/golden/, made by oracle/make_golden.py) plus the reference's own
hand-written golden vectors.  CPU only."""
import pytest
import torch

from conftest import bits_equal, load_golden
from oracle import ref_ops as R


def _ids(cases):
    return [f"{i}" for i in range(len(cases))]


STATIC = load_golden("static")


@pytest.mark.parametrize("i", range(len(STATIC)))
def test_static_ops_bit_exact(i):
    c = STATIC[i]
    q = R.quantize_by_tile(c["x"], c["scale"], c["tile"], float(c["num_bits"]), c["qdtype"], c["offset"])
    assert bits_equal(q, c["q"])
    y = R.dequantize_by_tile(q, c["scale"], c["tile"], c["offset"], c["ddtype"])
    assert bits_equal(y, c["y"])
    if "dx" in c:
        dx, dscale, doffset = R.quantize_by_tile_backward(
            c["x"], c["grad"], c["scale"], c["tile"], float(c["num_bits"]), c["offset"]
        )
        assert bits_equal(dx, c["dx"])
        # same aten sum on the same machine class: equal, but only a tolerance is contractual
        torch.testing.assert_close(dscale, c["dscale"], rtol=1e-5, atol=1e-5)
        if c["offset"] is not None:
            torch.testing.assert_close(doffset, c["doffset"], rtol=1e-5, atol=1e-5)
        else:
            assert doffset.numel() == 0


QUANTIZER = load_golden("quantizer")


@pytest.mark.parametrize("i", range(len(QUANTIZER)))
def test_linear_quantizer_flow(i):
    c = QUANTIZER[i]
    scale, offset = R.parameters_for_range(
        c["range_min"], c["range_max"], c["num_bits"], c["symmetric"], c["allow_one_sided"]
    )
    assert bits_equal(scale, c["scale"])
    stored = c["offset"]
    if stored is None:
        assert offset is None or True  # symmetric & !one-sided-allowed: no offset slot at all
        offset_used = None
    else:
        offset_used = offset if offset is not None else torch.zeros_like(scale)
        assert bits_equal(offset_used, stored)
    q = R.quantize_by_tile(c["x"], scale, c["tile"], c["num_bits"], c["x"].dtype, offset_used)
    assert bits_equal(q, c["q"])
    y = R.dequantize_by_tile(q, scale, c["tile"], offset_used, c["x"].dtype)
    assert bits_equal(y, c["y"])
    dx, dscale, doffset = R.quantize_by_tile_backward(c["x"], c["grad"], scale, c["tile"], c["num_bits"], offset_used)
    assert bits_equal(dx, c["dx"])
    torch.testing.assert_close(dscale, c["dscale"], rtol=1e-5, atol=1e-5)
    if c["doffset"] is not None:
        torch.testing.assert_close(doffset, c["doffset"], rtol=1e-5, atol=1e-5)
    lo, hi = R.quantization_range(scale, offset_used, c["num_bits"])
    assert bits_equal(lo, c["range_after"][0]) and bits_equal(hi, c["range_after"][1])


MINMAX = load_golden("running_minmax")


@pytest.mark.parametrize("i", range(len(MINMAX)))
def test_running_minmax(i):
    c = MINMAX[i]
    shape = c["shape"]
    tile = {
        "per_tensor": shape,
        "per_channel_last": (shape[0], shape[1], 1) if len(shape) == 3 else None,
        "per_channel0": (1,) + tuple(shape[1:]),
        "per_block": (1, 16),
    }[c["gran"]]
    mn = mx = None
    scale = offset = None
    for b in c["batches"]:
        # the estimator's state starts from the quantizer's current range once it has one
        # (minmax.py:190-200), which after the first step is the *representable* range
        mn, mx = R.running_minmax_step(mn, mx, b, tile)
        scale, offset = R.parameters_for_range(mn, mx, 8, c["symmetric"], True)
    assert bits_equal(scale, c["scale"])
    off = offset if offset is not None else torch.zeros_like(scale)
    assert bits_equal(off, c["offset"])
    last = c["batches"][-1]
    if c["disable_quantization"]:
        assert bits_equal(last, c["last_raw"])
    else:
        q = R.quantize_by_tile(last, scale, tile, 8, last.dtype, off)
        assert bits_equal(q, c["last_raw"])


CALIB8 = load_golden("calib_int8")


@pytest.mark.parametrize("i", range(len(CALIB8)))
def test_calibration_int8_recipe(i):
    """W8A8 calibration recipe (int8 codes): every step's codes, the final parameters and the running range."""
    c = CALIB8[i]
    shape = c["shape"]
    tile = shape if c["gran"] == "per_tensor" else (1,) + tuple(shape[1:])
    mn = mx = None
    for b, raw in zip(c["batches"], c["raws"]):
        mn, mx = R.running_minmax_step(mn, mx, b, tile)
        scale, offset = R.parameters_for_range(mn, mx, c["num_bits"], c["symmetric"], c["allow_one_sided"])
        q = R.quantize_by_tile(b, scale, tile, c["num_bits"], torch.int8, offset)
        assert bits_equal(q, raw)
    assert bits_equal(scale, c["scale"])
    if c["offset"] is None:
        assert offset is None
    else:
        assert bits_equal(offset if offset is not None else torch.zeros_like(scale), c["offset"])


GPTQ = load_golden("gptq")


@pytest.mark.parametrize("i", range(len(GPTQ)))
def test_gptq(i):
    """oracle/gptq_ref.py against the unmodified reference's gptq(): calibrated parameters, final weight and
    (for grouped quantizers) the recomputed group parameters, bit for bit."""
    from oracle import gptq_ref as G

    c = GPTQ[i]
    w = c["weight"]
    tile = c["tile"]
    # initial range estimation (gptq.py:76-77): smoothed_minmax of one batch == that batch's min/max
    mn, mx = R.smoothed_minmax_step(None, None, w.float(), tile, 1.0)
    scale, offset = R.parameters_for_range(mn, mx, c["num_bits"], c["symmetric"], True)
    assert bits_equal(scale, c["scale0"])
    if c["offset0"] is None:
        assert offset is None
    else:
        offset = offset if offset is not None else torch.zeros_like(scale)
        assert bits_equal(offset, c["offset0"])
    scale = scale.clone()
    offset = None if offset is None else offset.clone()
    # the trailing update between blocks is a library GEMM; its accumulation order depends on the thread count,
    # and the golden vectors were recorded single-threaded (oracle/make_golden.py)
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        new_w, _ = _run_gptq_oracle(G, c, w, scale, offset, tile)
    finally:
        torch.set_num_threads(threads)
    assert bits_equal(new_w, c["new_weight"])
    assert bits_equal(scale, c["scale"])
    if c["offset"] is not None:
        assert bits_equal(offset, c["offset"])


def _run_gptq_oracle(G, c, w, scale, offset, tile):
    return G.gptq(w, c["activations"], scale, offset, tile, c["num_bits"], c["symmetric"], True, c["qdtype"],
                  block_size=c["block_size"], actorder=c["actorder"], grouped=c["gran"] in ("per_block32", "per_tile"))


DYNAMIC = load_golden("dynamic")


@pytest.mark.parametrize("i", range(len(DYNAMIC)))
def test_dynamic(i):
    c = DYNAMIC[i]
    q, scale, offset = R.quantize_dynamic_by_tile(
        c["x"], c["tile"], float(c["num_bits"]), c["symmetric"], c["allow_one_sided"], c["qdtype"]
    )
    assert bits_equal(q, c["q"]) and bits_equal(scale, c["scale"]) and bits_equal(offset, c["offset"])


LINEAR = load_golden("linear")


@pytest.mark.parametrize("i", range(len(LINEAR)))
def test_fallback_linear(i):
    c = LINEAR[i]
    xd, wd = c["x"].dtype, c["w"].dtype
    y = R.fallback_linear(
        c["x_codes"], c["x_scale"], c["x_offset"], c["x"].shape, xd,
        c["w_codes"], c["w_scale"], c["w_offset"], (1, c["k"]), wd, c["bias"],
    )
    # the float GEMM's accumulation order is threading dependent: tolerance, not bits
    tol = dict(rtol=1e-5, atol=1e-5) if xd is torch.float32 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(y, c["y"], **tol)


# ---- the reference's own hand-written vectors --------------------------------------------
def test_ref_17_point_vector():
    # /root/reference/tests/nn/test_linear_quantizer.py:20-59 -- half-to-even + clamp, 2 bits
    data = torch.linspace(-8, 8, 17)
    scale = torch.tensor(2.0)
    q = R.quantize_by_tile(data, scale, data.shape, 2, None, None)
    expect = torch.tensor([-2.0] * 6 + [-1.0, 0.0, 0.0, 0.0] + [1.0] * 7)
    assert torch.equal(q, expect)
    assert torch.equal(R.dequantize_by_tile(q, scale, data.shape), expect * 2.0)


def test_ref_one_sided_offsets():
    # /root/reference/tests/nn/test_linear_quantizer.py:400-418
    mn, mx = torch.tensor([0.0]), torch.tensor([1.0])
    s, o = R.parameters_for_range(mn, mx, 4, symmetric=True, allow_one_sided=True)
    assert float(o) == 8.0
    s, o = R.parameters_for_range(mn, mx, 4, symmetric=True, allow_one_sided=False)
    assert o is None
    s, o = R.parameters_for_range(torch.tensor([-1.0]), mx, 4, symmetric=True, allow_one_sided=True)
    assert o is None


def test_ref_tile_order():
    # /root/reference/tests/quantization/test_tiled_tensor.py:10-43 -- tile index row-major over the grid
    data = torch.arange(24).reshape(4, 6)
    rows = R.tile_rows(data, (2, 3))
    assert rows.tolist() == [[0, 1, 2, 6, 7, 8], [3, 4, 5, 9, 10, 11], [12, 13, 14, 18, 19, 20], [15, 16, 17, 21, 22, 23]]
    assert torch.equal(R.untile_rows(rows, (4, 6), (2, 3)), data)
    with pytest.raises(ValueError):
        R.tile_rows(data, (3, 3))
    with pytest.raises(ValueError):
        R.tile_rows(data, (2,))


def test_bitwidth_guard():
    # _quantizer_impl.py:44-75,165-167
    assert R.can_support_bitwidth(torch.bfloat16, 8) and not R.can_support_bitwidth(torch.bfloat16, 16)
    assert R.can_support_bitwidth(torch.int8, 8) and R.can_support_bitwidth(torch.float16, 12)
    with pytest.raises(RuntimeError):
        R.quantize_by_tile(torch.randn(4), torch.tensor(1.0), (4,), 16, torch.bfloat16)


# ---- the plain-C restatement (oracle/ffq_oracle.c) against the same reference vectors -----------------
def test_c_oracle_matches_reference_vectors():
    from oracle import c_oracle as CO

    lib = CO.load()
    n = 0
    for c in STATIC:
        if not (c["x"].dtype is torch.float32 and c["scale"].dtype is torch.float32 and c["qdtype"] is torch.float32
                and (c["offset"] is None or c["offset"].dtype is torch.float32)):
            continue
        n += 1
        q = CO.quantize(lib, c["x"], c["scale"], c["tile"], c["num_bits"], c["offset"])
        assert bits_equal(q, c["q"])
        assert bits_equal(CO.dequantize(lib, q, c["scale"], c["tile"], c["offset"]), c["y"])
        dx, dscale, doffset = CO.backward(lib, c["x"], c["grad"], c["scale"], c["tile"], c["num_bits"], c["offset"])
        assert bits_equal(dx, c["dx"])
        torch.testing.assert_close(dscale.float().reshape(c["dscale"].shape), c["dscale"], rtol=1e-5, atol=1e-5)
        mn, mx = CO.minmax(lib, c["x"], c["tile"], c["scale"].numel())
        rmn, rmx = R.tile_minmax(c["x"], c["tile"])
        assert bits_equal(mn, rmn) and bits_equal(mx, rmx)
    assert n > 100
    for c in QUANTIZER:
        s, o = CO.params_for_range(lib, c["range_min"], c["range_max"], c["num_bits"], c["symmetric"], c["allow_one_sided"])
        assert bits_equal(s, c["scale"])
        if c["offset"] is not None:
            assert bits_equal(o if o is not None else torch.zeros_like(s), c["offset"])


MSE_GRID = load_golden("mse_grid")


def mse_grid_tile(c):
    shape = c["shape"]
    return {
        "per_tensor": tuple(shape),
        "per_channel_last": tuple(shape[:-1]) + (1,),
        "per_channel0": (1,) + tuple(shape[1:]),
        "per_block": (1, 16),
    }[c["gran"]]


@pytest.mark.parametrize("i", range(len(MSE_GRID)))
def test_mse_grid(i):
    """min_error.py:64-221: the search grid is bit-exact, the accumulated errors are the same aten
    mean on the same machine (tolerance is the contract), the selected range gives the recorded
    parameters."""
    c = MSE_GRID[i]
    tile = mse_grid_tile(c)
    batches = c["batches"]
    lo, hi = R.uniform_search_grid(batches[0], tile, c["symmetric"], c["num_candidates"])
    assert bits_equal(lo, c["min_threshold"]) and bits_equal(hi, c["max_threshold"])
    cumulative = torch.zeros_like(lo)
    for b in batches:
        cumulative += R.mse_grid_errors(b, tile, lo, hi, c["num_bits"], c["symmetric"], True,
                                        num_candidates=c["num_candidates"])
    torch.testing.assert_close(cumulative, c["cumulative_error"], rtol=1e-5, atol=1e-6)
    best_lo, best_hi = R.mse_grid_select(c["cumulative_error"], lo, hi)
    scale, offset = R.parameters_for_range(best_lo, best_hi, c["num_bits"], c["symmetric"], True)
    assert bits_equal(scale, c["scale"])
    assert bits_equal(offset if offset is not None else torch.zeros_like(scale), c["offset"])


def test_lpbq_oracle_matches_reference():
    """oracle.lpbq_grouped_dynamic_quantize against the reference's LPBQProcessor on both block orientations."""
    from conftest import load_golden
    cases = load_golden("lpbq")
    assert len(cases) == 15
    for c in cases:
        rows, cols = c["data_shape"]
        block = c["tile_size"][1] if c["orientation"] == "rows" else c["tile_size"][0]
        shape2d = (rows, cols // block) if c["orientation"] == "rows" else (rows // block, cols)
        q, f = R.lpbq_grouped_dynamic_quantize(c["scale"].reshape(shape2d), 0 if c["orientation"] == "rows" else 1,
                                               c["compressed_bw"])
        assert torch.equal(q, c["int_scale"]) and torch.equal(f.flatten(), c["float_scale"].flatten())
        assert c["encoding"]["per_block_int_scale"] == q.flatten().tolist()
        assert c["encoding"]["scale"] == f.flatten().tolist()
