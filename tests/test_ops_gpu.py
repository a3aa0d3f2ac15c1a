"""Parity of the CUDA ops (through the C ABI) against the reference: the golden vectors recorded
from the unmodified reference, and the CPU oracle on larger seeded inputs.

Bars (BASELINE.json north_star): integer codes, dequantized values, dx, min/max and derived
scale/offset are BIT-EXACT; per-tile gradient sums are compared with the float64-accumulated
oracle within  4 * eps * sum|terms|  (the reference's own aten sum is order dependent, SURVEY.md
section 7) and must be deterministic run to run."""
import pytest
import torch

from conftest import bits_equal, load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from fastforward_b200 import ops
    from oracle import ref_ops as R

DEV = "cuda"


def _cu(t):
    return None if t is None else t.to(DEV)


def _check_sums(got, data, grad, scale, tile, num_bits, offset, which):
    _, dsc, doff = R.backward_terms(data, grad, scale, tile, num_bits, offset)
    terms = dsc if which == "dscale" else doff
    exact = terms.double().sum(1)
    bound = terms.double().abs().sum(1)
    eps = torch.finfo(got.dtype).eps
    err = (got.detach().cpu().double().reshape(-1) - exact).abs()
    tol = 4 * eps * bound + 1e-30
    if got.dtype != torch.float32:      # result itself is rounded to a 16-bit type
        tol = tol + eps * exact.abs()
    assert bool((err <= tol).all()), f"{which}: max err {err.max()} tol {tol[err.argmax()]}"


STATIC = load_golden("static")


@pytest.mark.parametrize("i", range(len(STATIC)))
def test_golden_static(i):
    c = STATIC[i]
    x, s, o = _cu(c["x"]), _cu(c["scale"]), _cu(c["offset"])
    q = ops.quantize_by_tile(x, s, c["tile"], float(c["num_bits"]), c["qdtype"], o)
    assert bits_equal(q, c["q"])
    y = ops.dequantize_by_tile(q, s, c["tile"], o, c["ddtype"])
    assert bits_equal(y, c["y"])
    if c["x"].dtype.is_floating_point:
        yf, codes = ops.fake_quantize_by_tile(x, s, c["tile"], float(c["num_bits"]), c["qdtype"], o, c["ddtype"], return_codes=True)
        assert bits_equal(yf, c["y"]) and bits_equal(codes, c["q"])
    if "dx" in c:
        g = _cu(c["grad"])
        dx, dscale, doffset = ops.quantize_by_tile_backward(x, g, s, c["tile"], float(c["num_bits"]), o)
        assert bits_equal(dx, c["dx"])
        assert dscale.dtype == c["dscale"].dtype and dscale.shape == c["dscale"].shape
        _check_sums(dscale, c["x"], c["grad"], c["scale"], c["tile"], c["num_bits"], c["offset"], "dscale")
        if c["offset"] is not None:
            assert doffset.dtype == c["doffset"].dtype
            _check_sums(doffset, c["x"], c["grad"], c["scale"], c["tile"], c["num_bits"], c["offset"], "doffset")
        else:
            assert doffset.numel() == 0


DYNAMIC = load_golden("dynamic")


@pytest.mark.parametrize("i", range(len(DYNAMIC)))
def test_golden_dynamic(i):
    c = DYNAMIC[i]
    q, s, o = ops.quantize_dynamic_by_tile(_cu(c["x"]), c["tile"], float(c["num_bits"]), c["symmetric"],
                                           c["allow_one_sided"], c["qdtype"])
    assert bits_equal(s, c["scale"]) and bits_equal(o, c["offset"]) and bits_equal(q, c["q"])


QUANTIZER = load_golden("quantizer")


@pytest.mark.parametrize("i", range(len(QUANTIZER)))
def test_golden_params_for_range(i):
    c = QUANTIZER[i]
    n = c["scale"].numel()
    scale = torch.empty(n, device=DEV)
    offset = None if c["offset"] is None else torch.empty(n, device=DEV)
    ops.parameters_for_range_(_cu(c["range_min"]), _cu(c["range_max"]), c["num_bits"], c["symmetric"],
                              c["allow_one_sided"], scale, offset)
    assert bits_equal(scale, c["scale"])
    if offset is not None:
        assert bits_equal(offset, c["offset"])


MINMAX = load_golden("running_minmax")


@pytest.mark.parametrize("i", range(len(MINMAX)))
def test_golden_running_minmax(i):
    c = MINMAX[i]
    shape = c["shape"]
    tile = {"per_tensor": shape, "per_channel_last": (shape[0], shape[1], 1) if len(shape) == 3 else None,
            "per_channel0": (1,) + tuple(shape[1:]), "per_block": (1, 16)}[c["gran"]]
    n = c["scale"].numel()
    dt = c["batches"][0].dtype
    run_min = torch.full((n,), float("inf"), dtype=dt, device=DEV)
    run_max = torch.full((n,), float("-inf"), dtype=dt, device=DEV)
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    scale = torch.empty(n, device=DEV)
    offset = torch.empty(n, device=DEV)
    for b in c["batches"]:
        ops.running_minmax_update_(run_min, run_max, _cu(b), tile, flags)
        ops.parameters_for_range_(run_min, run_max, 8, c["symmetric"], True, scale, offset)
    assert int(flags.item()) == 0
    assert bits_equal(scale, c["scale"]) and bits_equal(offset, c["offset"])
    if not c["disable_quantization"]:
        last = _cu(c["batches"][-1])
        assert bits_equal(ops.quantize_by_tile(last, scale, tile, 8.0, last.dtype, offset), c["last_raw"])


# ---- oracle parity on larger seeded inputs --------------------------------------------------
BIG = [
    # shape, tile, dtype, num_bits, with_offset
    ((1024, 4096), (1, 4096), torch.float32, 8, True),      # cfg1 layout (rows reduced to keep the oracle fast)
    ((1024, 4096), (1, 4096), torch.bfloat16, 8, False),
    ((512, 4096), (1, 128), torch.float32, 4, True),        # cfg3 layout, g=128
    ((512, 4096), (1, 128), torch.bfloat16, 4, True),
    ((512, 4096), (1, 32), torch.float16, 4, True),
    ((2048, 4096), (2048, 4096), torch.float32, 8, True),   # per-tensor: segmented two-stage reduction
    ((2048, 4096), (2048, 4096), torch.bfloat16, 8, True),
    ((777, 1000), (777, 1000), torch.float32, 8, True),     # unaligned sizes, per tensor
    ((64, 1000), (1, 1000), torch.float32, 8, True),        # row length not a power of two
    ((300, 96), (1, 96), torch.bfloat16, 4, True),
    ((256, 512), (256, 1), torch.float32, 8, True),         # per-channel on the last dim (generic path)
    ((8, 128, 64), (8, 1, 64), torch.bfloat16, 8, True),    # per-channel(1) on 3-D
    ((16, 32, 48), (4, 8, 6), torch.float32, 3, True),      # arbitrary tiles
]


@pytest.mark.parametrize("shape,tile,dtype,num_bits,with_offset", BIG)
def test_oracle_parity_big(shape, tile, dtype, num_bits, with_offset):
    g = torch.Generator().manual_seed(hash((shape, tile, num_bits)) % (2 ** 31))
    x = (torch.randn(shape, generator=g) * 1.3).to(dtype)
    grad = torch.randn(shape, generator=g).to(dtype)
    rows = R.tile_rows(x.float(), tile)
    nt = rows.shape[0]
    rng = rows.abs().amax(1) * 0.7 + 1e-3            # 0.7: a share of elements clips
    scale = (rng / (2 ** (num_bits - 1))).float()
    offset = (torch.randn(nt, generator=g) * 2.0).float() if with_offset else None
    xc, gc, sc, oc = _cu(x), _cu(grad), _cu(scale), _cu(offset)

    q_ref = R.quantize_by_tile(x, scale, tile, num_bits, dtype, offset)
    q = ops.quantize_by_tile(xc, sc, tile, float(num_bits), dtype, oc)
    assert bits_equal(q, q_ref)
    q8 = ops.quantize_by_tile(xc, sc, tile, float(num_bits), torch.int8, oc)
    assert bits_equal(q8, R.quantize_by_tile(x, scale, tile, num_bits, torch.int8, offset))
    y_ref = R.dequantize_by_tile(q_ref, scale, tile, offset, dtype)
    assert bits_equal(ops.dequantize_by_tile(q, sc, tile, oc, dtype), y_ref)
    assert bits_equal(ops.dequantize_by_tile(q8, sc, tile, oc, dtype),
                      R.dequantize_by_tile(q8.cpu(), scale, tile, offset, dtype))
    assert bits_equal(ops.fake_quantize_by_tile(xc, sc, tile, float(num_bits), dtype, oc, dtype), y_ref)

    dx_ref, _, _ = R.quantize_by_tile_backward_f64(x, grad, scale, tile, num_bits, offset)
    dx, dscale, doffset = ops.quantize_by_tile_backward(xc, gc, sc, tile, float(num_bits), oc)
    assert bits_equal(dx, dx_ref)
    _check_sums(dscale, x, grad, scale, tile, num_bits, offset, "dscale")
    if with_offset:
        _check_sums(doffset, x, grad, scale, tile, num_bits, offset, "doffset")
    # deterministic: same bits on a second run
    dx2, dscale2, doffset2 = ops.quantize_by_tile_backward(xc, gc, sc, tile, float(num_bits), oc)
    assert bits_equal(dscale, dscale2) and bits_equal(doffset, doffset2) and bits_equal(dx, dx2)

    mn_ref, mx_ref = R.tile_minmax(x, tile)
    mn, mx = ops.tile_minmax(xc, tile)
    assert bits_equal(mn, mn_ref) and bits_equal(mx, mx_ref)

    qd, sd, od = ops.quantize_dynamic_by_tile(xc, tile, float(num_bits), False, True, dtype)
    qd_ref, sd_ref, od_ref = R.quantize_dynamic_by_tile(x, tile, float(num_bits), False, True, dtype)
    assert bits_equal(sd, sd_ref) and bits_equal(od, od_ref) and bits_equal(qd, qd_ref)


def test_edge_cases():
    # empty tensor, NaN propagation, inf flag, -0.0, bit-width guard, broadcast scale
    e = torch.empty(0, 4, device=DEV)
    assert ops.quantize_by_tile(e, torch.ones(1, device=DEV), (1, 4), 8.0, None).shape == (0, 4)
    x = torch.tensor([float("nan"), -0.3, 0.3, 1e9, -1e9, 2.5, 3.5, -2.5], device=DEV)
    s = torch.ones(1, device=DEV)
    q = ops.quantize_by_tile(x, s, (8,), 4.0, None)
    qr = R.quantize_by_tile(x.cpu(), s.cpu(), (8,), 4.0, None)
    assert bits_equal(q, qr)                      # includes NaN passthrough and -0.0
    mn, mx = ops.tile_minmax(x, (8,))
    assert torch.isnan(mn).all() and torch.isnan(mx).all()
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    rmin = torch.full((1,), float("inf"), device=DEV)
    rmax = torch.full((1,), float("-inf"), device=DEV)
    ops.running_minmax_update_(rmin, rmax, torch.tensor([1.0, float("inf")], device=DEV), (2,), flags)
    assert int(flags.item()) == 1
    with pytest.raises(RuntimeError):
        ops.quantize_by_tile(x, s, (8,), 16.0, torch.bfloat16)
    with pytest.raises(ValueError):
        ops.quantize_by_tile(x, s, (3,), 4.0, None)
    with pytest.raises(ValueError):
        ops.quantize_by_tile(x, s, (2, 4), 4.0, None)
    with pytest.raises(RuntimeError):
        ops.quantize_by_tile(x, torch.ones(3, device=DEV), (2,), 4.0, None)   # 3 params for 4 tiles
    # a one-element scale broadcasts over all tiles, like scale[:, None] does
    q2 = ops.quantize_by_tile(x, s, (2,), 4.0, None)
    assert bits_equal(q2, R.quantize_by_tile(x.cpu(), s.cpu().expand(4), (2,), 4.0, None))


def test_unaligned_view():
    base = torch.randn(4097, device=DEV)
    x = base[1:]                                   # 4-byte aligned only
    s = torch.full((1,), 0.05, device=DEV)
    assert bits_equal(ops.quantize_by_tile(x, s, (4096,), 8.0, None),
                      R.quantize_by_tile(x.cpu(), s.cpu(), (4096,), 8.0, None))
    g = torch.randn(4097, device=DEV)[1:]
    dx, dsc, _ = ops.quantize_by_tile_backward(x, g, s, (4096,), 8.0, None)
    assert bits_equal(dx, R.quantize_by_tile_backward(x.cpu(), g.cpu(), s.cpu(), (4096,), 8.0, None)[0])


def test_full_size_properties():
    """BASELINE cfg1 at full size (4096x4096 fp32, 8-bit per-channel): size-independent properties."""
    torch.manual_seed(0)
    x = torch.randn(4096, 4096, device=DEV)
    g = torch.randn(4096, 4096, device=DEV)
    tile = (1, 4096)
    mn, mx = ops.tile_minmax(x, tile)
    assert torch.equal(mn, x.min(1).values) and torch.equal(mx, x.max(1).values)
    scale = torch.empty(4096, device=DEV)
    offset = torch.empty(4096, device=DEV)
    ops.parameters_for_range_(mn, mx, 8, True, True, scale, offset)
    q = ops.quantize_by_tile(x, scale, tile, 8.0, torch.int8, offset)
    assert int(q.max()) <= 127 and int(q.min()) >= -128
    y = ops.dequantize_by_tile(q, scale, tile, offset, torch.float32)
    assert float((y - x).abs().max()) <= float(scale.max()) * 0.5 * (1 + 1e-6)   # no clipping inside the range
    # idempotence: fake-quant of a fake-quantized tensor is itself (fuse.py relies on it, test_fuse.py:51-117)
    y1 = ops.fake_quantize_by_tile(x, scale, tile, 8.0, None, offset)
    assert torch.equal(y1, y)
    assert torch.equal(ops.fake_quantize_by_tile(y1, scale, tile, 8.0, None, offset), y1)
    # linearity of the backward in g, and dx == g where nothing clips
    dx, dsc, doff = ops.quantize_by_tile_backward(x, g, scale, tile, 8.0, offset)
    assert torch.equal(dx, g)
    dx2, dsc2, doff2 = ops.quantize_by_tile_backward(x, 2 * g, scale, tile, 8.0, offset)
    assert torch.equal(dx2, 2 * g) and torch.equal(dsc2, 2 * dsc) and torch.equal(doff2, 2 * doff)


def test_shared_reciprocal_division_is_ieee_exact():
    """The kernels hoist the reciprocal out of the per-element work (ffq_common.cuh: shared_div).
    Inside its guard the quotient must equal __fdiv_rn bit for bit: sweep 2^30 pairs."""
    from fastforward_b200 import _cabi as C
    counts = torch.zeros(4, dtype=torch.int64, device=DEV)
    C.check(C.lib.ffq_selftest_shared_div(1 << 30, 12345, counts.data_ptr(), C.current_stream(counts.device)))
    bad, acc, acc_strict, bad_strict = counts.tolist()
    assert bad == 0 and bad_strict == 0, counts.tolist()
    assert acc > (1 << 28) and acc_strict > (1 << 28), counts.tolist()   # the guard accepts the bulk


@pytest.mark.parametrize("rows", [512, 4104])
def test_host_buffer_entry_point(rows):
    """ffq_fakequant_fwd_bwd_host: the end-to-end C-ABI call with HOST buffers (H2D, two kernels, D2H).  4104 rows
    (16 MB) take the pipelined path: eight groups of whole tiles, the last one ragged, copies overlapping the kernels."""
    import ctypes
    from fastforward_b200 import _cabi as C

    torch.manual_seed(11)
    shape, tile = (rows, 1024), (1, 1024)
    x, g = torch.randn(shape), torch.randn(shape)
    scale = (x.abs().amax(1) * 0.6 / 128).contiguous()
    offset = (torch.randn(shape[0]) * 3).contiguous()
    y, dx = torch.empty_like(x), torch.empty_like(x)
    dsc, dof = torch.empty(shape[0]), torch.empty(shape[0])
    lay = C.make_layout(shape, tile)
    C.check(C.lib.ffq_fakequant_fwd_bwd_host(x.data_ptr(), g.data_ptr(), C.dtype_tag(x.dtype), y.data_ptr(), dx.data_ptr(),
                                              dsc.data_ptr(), dof.data_ptr(), scale.data_ptr(), offset.data_ptr(),
                                              ctypes.byref(lay), 8.0, 0))
    q = R.quantize_by_tile(x, scale, tile, 8, x.dtype, offset)
    assert bits_equal(y, R.dequantize_by_tile(q, scale, tile, offset, x.dtype))
    rdx, rdsc, rdof = R.quantize_by_tile_backward_f64(x, g, scale, tile, 8, offset)
    assert bits_equal(dx, rdx)
    # per-tile sums of 1024 fp32 terms against the fp64 oracle (contract: 4 eps sum|terms|; |terms| <= ~4 here)
    torch.testing.assert_close(dsc.double(), rdsc, rtol=1e-4, atol=5e-4)
    torch.testing.assert_close(dof.double(), rdof, rtol=1e-4, atol=5e-4)
