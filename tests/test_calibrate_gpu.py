"""Fused calibration step (ffq_calibrate_quantize): min/max + running-range update + range->params + int8
quantize (+ code row sums) in one kernel.  Checked bit for bit against (1) vectors recorded from the
unmodified reference running the W8A8 calibration recipe (tests/golden/calib_int8.pt.gz), (2) the CPU oracle
and (3) the unfused kernel sequence, including the deferred one-sided rows, NaN/inf data, ragged sizes,
multi-chunk tensors and CUDA-graph replay."""
import pytest
import torch

from conftest import bits_equal, load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import fastforward_b200 as ff
    from fastforward_b200 import ops
    from oracle import ref_ops as R

DEV = "cuda"
CALIB8 = load_golden("calib_int8")


def _gran(name):
    return ff.PerTensor() if name == "per_tensor" else ff.PerChannel(0)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("i", range(len(CALIB8)))
def test_estimate_ranges_int8_matches_reference(i, fused):
    c = CALIB8[i]
    q = ff.nn.LinearQuantizer(c["num_bits"], symmetric=c["symmetric"], allow_one_sided=c["allow_one_sided"],
                              granularity=_gran(c["gran"]), quantized_dtype=torch.int8, device=DEV)
    launches = []
    with torch.no_grad(), ff.estimate_ranges(q, ff.range_setting.running_minmax, fused=fused):
        for b, raw in zip(c["batches"], c["raws"]):
            before = ff._cabi.launch_count()
            out = q(b.to(DEV))
            launches.append(ff._cabi.launch_count() - before)
            assert isinstance(out, ff.QuantizedTensor) and out.raw_data.dtype == torch.int8
            assert bits_equal(out.raw_data, raw)
            assert out.quant_args().dequantize_dtype == b.dtype
            if fused:
                assert out._ffq_rowsum.cpu().tolist() == raw.int().sum(-1).reshape(-1).tolist()
    assert bits_equal(q.scale.detach(), c["scale"])
    if c["offset"] is None:
        assert q.offset is None
    else:
        assert bits_equal(q.offset.detach(), c["offset"])
    lo, hi = q.quantization_range
    assert bits_equal(lo.detach(), c["range"][0]) and bits_equal(hi.detach(), c["range"][1])
    if fused:   # one kernel (+ the one-sided fix-up pass behind per-channel symmetric quantizers)
        rows_fixup = c["gran"] == "per_channel0" and c["symmetric"] and c["allow_one_sided"]
        assert max(launches) <= (2 if rows_fixup else 1)
    else:
        assert min(launches) >= 3


def _oracle_step(mn, mx, x, tile, bits, symmetric, one_sided):
    mn, mx = R.running_minmax_step(mn, mx, x, tile)
    scale, offset = R.parameters_for_range(mn, mx, bits, symmetric, one_sided)
    q = R.quantize_by_tile(x, scale, tile, bits, torch.int8, offset)
    return mn, mx, scale, offset, q


def _run_fused(xs, tile, bits, symmetric, one_sided, rowsum=True, run_dtype=None):
    x0 = xs[0]
    nt = 1
    for d, t in zip(x0.shape, tile):
        nt *= d // t
    rdt = run_dtype or x0.dtype
    mn = torch.full((nt,), float("inf"), dtype=rdt, device=DEV)
    mx = torch.full((nt,), float("-inf"), dtype=rdt, device=DEV)
    scale = torch.empty(nt, device=DEV)
    offset = None if (symmetric and not one_sided) else torch.empty(nt, device=DEV)
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    settled = torch.zeros(1, dtype=torch.int32, device=DEV)
    outs = []
    for x in xs:
        q, rs = ops.calibrate_quantize_(mn, mx, x.to(DEV), tile, bits, symmetric, one_sided, scale, offset, flags, settled,
                                        rowsum=rowsum)
        outs.append((q.cpu(), None if rs is None else rs.cpu(), scale.cpu().clone(), None if offset is None else offset.cpu().clone(),
                     mn.cpu().clone(), mx.cpu().clone()))
    return outs, int(flags.item()), int(settled.item())


def _check_against_oracle(xs, tile, bits, symmetric, one_sided):
    outs, flags, _ = _run_fused(xs, tile, bits, symmetric, one_sided)
    mn = mx = None
    for x, (q, rs, scale, offset, rmn, rmx) in zip(xs, outs):
        mn, mx, s, o, rq = _oracle_step(mn, mx, x, tile, bits, symmetric, one_sided)
        assert bits_equal(rmn, mn) and bits_equal(rmx, mx)
        assert bits_equal(scale, s)
        if offset is not None:
            assert bits_equal(offset, o if o is not None else torch.zeros_like(s))
        assert bits_equal(q, rq)
        assert torch.equal(rs, rq.reshape(-1, rq.shape[-1]).int().sum(1).to(torch.int32))
    return flags


ROW_SHAPES = [((7, 512), torch.bfloat16), ((3, 4096), torch.bfloat16), ((5, 14336), torch.bfloat16), ((4, 1000 * 8), torch.float16),
              ((9, 256), torch.float32), ((2, 16384), torch.float32), ((3, 5000), torch.float32), ((301, 1024), torch.bfloat16)]


@pytest.mark.parametrize("shape,dtype", ROW_SHAPES)
@pytest.mark.parametrize("symmetric,one_sided", [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize("variant", ["mixed", "positive", "some_positive"])
def test_rows_vs_oracle(shape, dtype, symmetric, one_sided, variant):
    g = torch.Generator().manual_seed(sum(shape) * 8 + symmetric * 4 + one_sided * 2 + len(variant))
    xs = []
    for i in range(3):
        x = torch.randn(shape, generator=g) * (0.02 + 0.5 * i)
        if variant == "positive":
            x = x.abs() + 1e-3
        elif variant == "some_positive":
            x[::2] = x[::2].abs()
        xs.append(x.to(dtype))
    assert _check_against_oracle(xs, (1, shape[1]), 8, symmetric, one_sided) == 0


TENSOR_SHAPES = [((8, 512), torch.bfloat16), ((2048, 4096), torch.bfloat16), ((1500, 14336), torch.bfloat16), ((3, 77, 264), torch.float16),
                 ((64, 256), torch.float32), ((1024, 4096), torch.float32), ((2, 5, 104), torch.float32)]


@pytest.mark.parametrize("shape,dtype", TENSOR_SHAPES)
@pytest.mark.parametrize("symmetric,one_sided", [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize("variant", ["mixed", "positive"])
def test_tensor_vs_oracle(shape, dtype, symmetric, one_sided, variant):
    g = torch.Generator().manual_seed(sum(shape) * 8 + symmetric * 4 + one_sided * 2 + len(variant))
    xs = []
    for i in range(2):
        x = torch.randn(shape, generator=g) * (0.3 + 1.5 * i) + 0.2
        if variant == "positive":
            x = x.abs()
        xs.append(x.to(dtype))
    assert _check_against_oracle(xs, tuple(shape), 8, symmetric, one_sided) == 0


OPT_SHAPES = [((512, 1024), torch.bfloat16), ((2048, 4096), torch.bfloat16), ((300, 14336), torch.bfloat16), ((4, 128, 1024), torch.float16),
              ((1000, 1004), torch.float32), ((64, 65536), torch.bfloat16), ((3, 70000 * 8), torch.bfloat16), ((40000, 8), torch.float32)]


@pytest.mark.parametrize("shape,dtype", OPT_SHAPES)
@pytest.mark.parametrize("symmetric,one_sided,bits", [(True, True, 8), (True, False, 8), (False, True, 8), (False, True, 4)])
@pytest.mark.parametrize("rowsum", [True, False])
def test_tensor_sequences_that_do_and_do_not_widen_the_range(shape, dtype, symmetric, one_sided, bits, rowsum):
    """The optimistic per-tensor pair (codes speculatively from the old range, rewritten only when the batch widens it)
    over a sequence of batches: fresh range, the same batch again, a narrower batch, a wider one, a narrower one again,
    a batch that moves only the minimum.  Every step's range, parameters, codes and row sums equal the oracle's.  The
    last three shapes have rows outside the pair's row-sum window and stay on the two-pass kernel."""
    g = torch.Generator().manual_seed(sum(shape) + symmetric * 4 + one_sided * 2 + bits)
    x0 = torch.randn(shape, generator=g) * 0.5 + 0.1
    x1 = torch.randn(shape, generator=g) * 2.0
    low = x0.clone()
    low.view(-1)[17] = -9.0
    xs = [t.to(dtype) for t in (x0, x0, x0 * 0.5, x1, x1 * 0.9, low, x0)]
    tile = tuple(shape)
    outs, flags, _ = _run_fused(xs, tile, bits, symmetric, one_sided, rowsum=rowsum)
    assert flags == 0
    mn = mx = None
    for i, (x, (q, rs, scale, offset, rmn, rmx)) in enumerate(zip(xs, outs)):
        mn, mx, s, o, rq = _oracle_step(mn, mx, x, tile, bits, symmetric, one_sided)
        assert bits_equal(rmn, mn) and bits_equal(rmx, mx), i
        assert bits_equal(scale, s), i
        if offset is not None:
            assert bits_equal(offset, o if o is not None else torch.zeros_like(s)), i
        assert bits_equal(q, rq), i
        if rowsum:
            assert torch.equal(rs, rq.reshape(-1, rq.shape[-1]).int().sum(1).to(torch.int32)), i
        else:
            assert rs is None


def test_tensor_sequence_with_special_values_and_stale_parameters():
    """NaN and +-inf inside a sequence (the range turns NaN / infinite and stays so: every later batch is rewritten), and
    parameters that do not match the running range when a step starts (the pair derives its own from the range)."""
    g = torch.Generator().manual_seed(3)
    shape = (256, 2048)
    x = (torch.randn(shape, generator=g)).bfloat16()
    bad = x.clone(); bad[7, 9] = float("nan")
    inf = x.clone(); inf[1, 1] = float("inf")
    for seq, want_flag in (([x, x, bad, x], 0), ([x, inf, x], 1)):
        outs, flags, _ = _run_fused(seq, shape, 8, False, True)
        assert (flags & 1) == want_flag
        mn = mx = None
        for i, (b, (q, rs, scale, offset, rmn, rmx)) in enumerate(zip(seq, outs)):
            if bool(torch.isinf(b.float()).any()):
                break                       # the reference raises at this batch (minmax.py:233-234): ours sets the flag
            mn, mx, s, o, rq = _oracle_step(mn, mx, b, shape, 8, False, True)
            assert bits_equal(rmn, mn) and bits_equal(rmx, mx) and bits_equal(scale, s) and bits_equal(offset, o), i
            assert bits_equal(q, rq), i
    # stale parameters: overwrite scale/offset between two steps on the same data
    mn = torch.full((1,), float("inf"), dtype=torch.bfloat16, device=DEV); mx = -mn
    scale, offset = torch.empty(1, device=DEV), torch.empty(1, device=DEV)
    xd = x.to(DEV)
    q1, rs1 = ops.calibrate_quantize_(mn, mx, xd, shape, 8, False, True, scale, offset, rowsum=True)
    s1, o1 = scale.clone(), offset.clone()
    scale.fill_(123.0); offset.fill_(-7.0)
    q2, rs2 = ops.calibrate_quantize_(mn, mx, xd, shape, 8, False, True, scale, offset, rowsum=True)
    assert torch.equal(q1, q2) and torch.equal(rs1, rs2) and torch.equal(scale, s1) and torch.equal(offset, o1)


def test_tensor_pair_under_graph_replay():
    """Captured once, replayed over changing inputs: the rewrite kernel's work depends on the data, its launch does not."""
    g = torch.Generator().manual_seed(9)
    shape = (1024, 1024)
    xs = [(torch.randn(shape, generator=g) * sc).bfloat16() for sc in (1.0, 0.5, 3.0, 0.1, 3.5)]
    mn = torch.full((1,), float("inf"), dtype=torch.bfloat16, device=DEV); mx = -mn
    scale, offset = torch.empty(1, device=DEV), torch.empty(1, device=DEV)
    ws = torch.zeros(ops._CALQ_WS, dtype=torch.uint8, device=DEV)
    static = xs[0].to(DEV).clone()
    ops.calibrate_quantize_(mn, mx, static, shape, 8, False, True, scale, offset, rowsum=True, workspace=ws)   # warm-up
    mn.fill_(float("inf")); mx.fill_(float("-inf"))
    torch.cuda.synchronize()
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        q, rs = ops.calibrate_quantize_(mn, mx, static, shape, 8, False, True, scale, offset, rowsum=True, workspace=ws)
    omn = omx = None
    for x in xs:
        static.copy_(x)
        cg.replay()
        omn, omx, s, o, rq = _oracle_step(omn, omx, x, shape, 8, False, True)
        assert bits_equal(q.cpu(), rq) and bits_equal(scale.cpu(), s) and bits_equal(offset.cpu(), o)
        assert torch.equal(rs.cpu(), rq.int().sum(1).to(torch.int32))


@pytest.mark.parametrize("bits", [2, 4, 7])
def test_low_bit_widths(bits):
    g = torch.Generator().manual_seed(bits)
    xs = [torch.randn(6, 1024, generator=g).bfloat16() for _ in range(2)]
    _check_against_oracle(xs, (1, 1024), bits, True, True)
    _check_against_oracle(xs, (6, 1024), bits, False, True)


def test_special_values():
    """NaN propagates into the range and the parameters exactly as torch.min/max do; +-inf sets the flag;
    huge / tiny magnitudes take the exact-division path."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 1024, generator=g)
    x[1, 5] = float("nan")
    x[2, 7] = 3e38
    x[3, :] = 0.0
    x[4, 9] = -1e-30
    for tile in [(1, 1024), (6, 1024)]:
        for symmetric in (True, False):
            assert _check_against_oracle([x, x * 0.5], tile, 8, symmetric, True) == 0
    x[5, 1] = float("inf")
    outs, flags, _ = _run_fused([x], (1, 1024), 8, False, True)
    assert flags & 1
    x[1, 5] = 0.0          # a NaN extremum is not infinite: the reference does not raise for it either
    outs, flags, _ = _run_fused([x], (6, 1024), 8, False, True)
    assert flags & 1


def test_settled_flag_and_range_dtype():
    """`settled` latches once every running min is negative; an fp32 running range continued with bf16 data."""
    g = torch.Generator().manual_seed(11)
    xs = [torch.randn(8, 2048, generator=g).bfloat16() for _ in range(3)]
    outs, _, settled = _run_fused(xs, (1, 2048), 8, True, True, run_dtype=torch.float32)
    assert settled == 1
    mn = mx = None
    for x, (q, rs, scale, offset, rmn, rmx) in zip(xs, outs):
        mn, mx, s, o, rq = _oracle_step(mn, mx, x, (1, 2048), 8, True, True)
        assert bits_equal(rmn, mn.float()) and bits_equal(scale, s) and bits_equal(q, rq)
    pos = [x.abs() for x in xs]
    _, _, settled = _run_fused(pos, (1, 2048), 8, True, True)
    assert settled == 0


def test_matches_unfused_kernels_large():
    """Llama-3-8B weight / activation shapes: identical bits to minmax + params_for_range + quantize + rowsum."""
    torch.manual_seed(3)
    for shape, tile, symmetric in [((4096, 4096), (1, 4096), True), ((14336, 4096), (1, 4096), True), ((4096, 14336), (1, 14336), True),
                                   ((2048, 4096), (2048, 4096), False), ((2048, 14336), (2048, 14336), False)]:
        x = (torch.randn(shape, device=DEV) * 0.02).bfloat16()
        nt = (shape[0] // tile[0]) * (shape[1] // tile[1])
        mn = torch.full((nt,), float("inf"), dtype=torch.bfloat16, device=DEV); mx = -mn
        mn2, mx2 = mn.clone(), mx.clone()
        scale, offset = torch.empty(nt, device=DEV), torch.empty(nt, device=DEV)
        scale2, offset2 = torch.empty(nt, device=DEV), torch.empty(nt, device=DEV)
        q, rs = ops.calibrate_quantize_(mn, mx, x, tile, 8, symmetric, True, scale, offset, rowsum=True)
        ops.running_minmax_update_(mn2, mx2, x, tile)
        ops.parameters_for_range_(mn2, mx2, 8, symmetric, True, scale2, offset2)
        q2 = ops.quantize_by_tile(x, scale2, tile, 8.0, torch.int8, offset2)
        assert torch.equal(mn, mn2) and torch.equal(mx, mx2) and torch.equal(scale, scale2) and torch.equal(offset, offset2)
        assert torch.equal(q, q2)
        assert torch.equal(rs, q2.int().sum(1).to(torch.int32))


def test_unsupported_layouts_fall_back():
    assert ops.calibrate_quantize_mode((64, 384), (1, 384), torch.bfloat16) == 0       # 48 vectors: neither a group nor a row
    assert ops.calibrate_quantize_mode((64, 128), (1, 128), torch.bfloat16) == 3       # per-group tiles
    assert ops.calibrate_quantize_mode((64, 4096), (64, 1), torch.bfloat16) == 0       # strided tiles
    assert ops.calibrate_quantize_mode((64, 4096), (1, 4096), torch.int32) == 0
    assert ops.calibrate_quantize_mode((64, 4096), (1, 4096), torch.bfloat16) == 1
    assert ops.calibrate_quantize_mode((64, 4096), (64, 4096), torch.bfloat16) == 2
    # the estimator silently takes the separate kernels there: odd tile sizes, float codes, grad mode
    q = ff.nn.LinearQuantizer(8, granularity=ff.PerBlock(block_dims=1, block_sizes=96, per_channel_dims=0),
                              quantized_dtype=torch.int8, device=DEV)
    x = torch.randn(16, 576, device=DEV)
    with torch.no_grad(), ff.estimate_ranges(q, ff.range_setting.running_minmax):
        out = q(x)
    assert not hasattr(out, "_ffq_rowsum")
    q2 = ff.nn.LinearQuantizer(8, granularity=ff.PerChannel(0), device=DEV)             # float codes
    with torch.no_grad(), ff.estimate_ranges(q2, ff.range_setting.running_minmax):
        out = q2(x)
    assert out.raw_data.dtype == torch.float32 and not hasattr(out, "_ffq_rowsum")


def test_quantized_linear_calibration_fused_equals_unfused_and_graph_replay():
    """A QuantizedLinear calibrated with the fused step gives the same parameters and outputs as with the separate
    kernels, consumes the attached row sums, and the whole step (cooperative kernel included) replays from a CUDA graph."""
    from fastforward_b200.nn import qlinear
    qlinear.install()
    torch.manual_seed(0)

    def build():
        torch.manual_seed(0)
        lin = torch.nn.Linear(1024, 768, bias=True).to(DEV).bfloat16()
        m = torch.nn.Sequential(lin)
        ff.quantize_model(m)
        ff.find_quantizers(m, "**/[quantizer:parameter/weight]").initialize(
            ff.nn.LinearQuantizer, num_bits=8, granularity=ff.PerChannel(0), quantized_dtype=torch.int8)
        ff.find_quantizers(m, "**/[quantizer:activation/input]").initialize(
            ff.nn.LinearQuantizer, num_bits=8, symmetric=False, granularity=ff.PerTensor(), quantized_dtype=torch.int8)
        return m.to(DEV)

    xs = [(torch.randn(2, 96, 1024, device=DEV) * (1 + i)).bfloat16() for i in range(3)]
    results = {}
    for fused in (True, False):
        m = build()
        ys = []
        before = ff._cabi.launch_count()
        with torch.no_grad(), ff.estimate_ranges(m, ff.range_setting.running_minmax, fused=fused):
            for x in xs:
                ys.append(m(x))
        results[fused] = (ys, [p.detach().clone() for _, q in ff.nn.named_quantizers(m) for p in (q.scale, q.offset)],
                          ff._cabi.launch_count() - before)
    for a, b in zip(results[True][0], results[False][0]):
        assert torch.equal(a, b)
    for a, b in zip(results[True][1], results[False][1]):
        assert torch.equal(a, b)
    assert results[True][2] < results[False][2]

    m = build()
    static = xs[0].clone()
    with torch.no_grad(), ff.estimate_ranges(m, ff.range_setting.running_minmax):
        for _ in range(2):
            m(static)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = m(static)
        for x, want in zip(xs, results[True][0]):
            static.copy_(x)
            g.replay()
        torch.cuda.synchronize()
    # the replays saw xs[0] (x3 incl. warm-up and capture), xs[0..2]: same running range as the eager run
    assert torch.equal(y, results[True][0][2])
    for (_, q), want in zip(ff.nn.named_quantizers(m), zip(results[True][1][0::2], results[True][1][1::2])):
        assert torch.equal(q.scale.detach(), want[0]) and torch.equal(q.offset.detach(), want[1])


def test_batched_parameters_for_ranges_equals_per_quantizer_calls():
    """The block-exit path of data-parallel calibration: one launch for all quantizers of an arena chunk."""
    g = torch.Generator().manual_seed(21)
    sizes = [1, 4096, 1, 14336, 7, 1, 300]
    cfgs = [(8, False, True), (8, True, True), (4, True, False), (8, True, True), (2, False, True), (8, True, True), (6, True, True)]
    total = sum(sizes)
    for dt in (torch.bfloat16, torch.float32):
        mn = (torch.randn(total, generator=g) - 0.5).to(dt).to(DEV)
        mx = (mn.float() + torch.rand(total, generator=g).to(DEV) * 3).to(dt)
        mn[sizes[0]:sizes[0] + 4096] = mn[sizes[0]:sizes[0] + 4096].abs()        # a one-sided quantizer
        mn[-300] = float("nan")                                                    # NaN range: never one-sided
        entries, expect, start = [], [], 0
        for n, (bits, sym, one_sided) in zip(sizes, cfgs):
            scale = torch.empty(n, device=DEV)
            offset = None if (sym and not one_sided) else torch.empty(n, device=DEV)
            entries.append((start, n, bits, sym, one_sided, scale, offset))
            s2 = torch.empty(n, device=DEV)
            o2 = None if offset is None else torch.empty(n, device=DEV)
            ops.parameters_for_range_(mn[start:start + n], mx[start:start + n], bits, sym, one_sided, s2, o2)
            expect.append((s2, o2))
            start += n
        before = ff._cabi.launch_count()
        ops.parameters_for_ranges_batched_(mn, mx, entries)
        assert ff._cabi.launch_count() - before == 1
        for (_, _, _, _, _, scale, offset), (s2, o2) in zip(entries, expect):
            assert bits_equal(scale, s2)
            assert (offset is None) == (o2 is None) and (offset is None or bits_equal(offset, o2))


def _unfused(run_mn, run_mx, x, tile, bits, sym, one_sided, flags):
    nt = run_mn.numel()
    scale = torch.empty(nt, device=DEV)
    offset = None if (sym and not one_sided) else torch.empty(nt, device=DEV)
    ops.running_minmax_update_(run_mn, run_mx, x, tile, flags)
    ops.parameters_for_range_(run_mn, run_mx, bits, sym, one_sided, scale, offset)
    q = ops.quantize_by_tile(x, scale, tile, float(bits), torch.int8, offset)
    return q, scale, offset


@pytest.mark.parametrize("seed", range(40))
def test_fused_equals_separate_kernels_on_random_configurations(seed):
    """Property test at sizes the CPU oracle would be slow for: for random shapes / dtypes / bit widths / quantizer
    kinds / data scales (with NaN, inf, zeros and huge values sprinkled in) three consecutive fused steps produce
    bit for bit what minmax + params_for_range + quantize (+ row sums) produce."""
    import random
    rnd = random.Random(1000 + seed)
    g = torch.Generator().manual_seed(seed)
    dt = rnd.choice([torch.bfloat16, torch.float16, torch.float32])
    ept = 4 if dt == torch.float32 else 8
    per_tensor = rnd.random() < 0.4
    if per_tensor:
        shape = rnd.choice([(rnd.randint(1, 900), ept * rnd.randint(8, 700)), (rnd.randint(1, 5), rnd.randint(1, 60), ept * rnd.randint(8, 300))])
        tile = shape
        if shape[-1] * (shape[0] if len(shape) == 2 else shape[0] * shape[1]) < 64 * ept:
            shape = (8, 64 * ept); tile = shape
    else:
        shape = (rnd.randint(1, 400), ept * rnd.randint(64, 2300))
        tile = (1, shape[1])
    bits = rnd.choice([8, 8, 8, 4, 3, 7])
    sym = rnd.random() < 0.5
    one_sided = rnd.random() < 0.7 or not sym
    nt = 1 if per_tensor else shape[0]
    rdt = rnd.choice([dt, torch.float32])
    mn1 = torch.full((nt,), float("inf"), dtype=rdt, device=DEV); mx1 = -mn1
    mn2, mx2 = mn1.clone(), mx1.clone()
    scale = torch.empty(nt, device=DEV)
    offset = None if (sym and not one_sided) else torch.empty(nt, device=DEV)
    f1 = torch.zeros(1, dtype=torch.int32, device=DEV); f2 = torch.zeros(1, dtype=torch.int32, device=DEV)
    settled = torch.zeros(1, dtype=torch.int32, device=DEV)
    special = rnd.random() < 0.35
    for step in range(3):
        x = torch.randn(shape, generator=g) * (10 ** rnd.uniform(-3, 2))
        kind = rnd.random()
        if kind < 0.2:
            x = x.abs()
        elif kind < 0.35 and not per_tensor:
            x[::3] = x[::3].abs()
        if special:
            flat = x.view(-1)
            for v in (float("nan"), float("inf"), 0.0, -0.0, 3e38, -1e-38):
                if rnd.random() < 0.4:
                    flat[rnd.randrange(flat.numel())] = v
        x = x.to(dt).to(DEV)
        q, rs = ops.calibrate_quantize_(mn1, mx1, x, tile, bits, sym, one_sided, scale, offset, f1, settled, rowsum=True)
        q2, s2, o2 = _unfused(mn2, mx2, x, tile, bits, sym, one_sided, f2)
        assert bits_equal(mn1, mn2) and bits_equal(mx1, mx2), "running range"
        assert bits_equal(scale, s2), "scale"
        assert offset is None or bits_equal(offset, o2), "offset"
        assert torch.equal(q, q2), "codes"
        assert torch.equal(rs, q2.reshape(-1, shape[-1]).int().sum(1).to(torch.int32)), "row sums"
        assert int(f1.item()) == int(f2.item())


# ---------------------------------------------------------------------------------------------------------------
# per-group tiles (mode 3) and the fused calibrate + fake-quantize (weight QDQ) entry point
# ---------------------------------------------------------------------------------------------------------------
def _variants(shape, g, variant, dt, steps):
    xs = []
    for i in range(steps):
        x = torch.randn(shape, generator=g) * (0.02 + 0.3 * i)
        if variant == "positive":
            x = x.abs() + 1e-3
        elif variant == "some_positive":
            x[::2] = x[::2].abs()
        xs.append(x.to(dt))
    return xs


@pytest.mark.parametrize("shape,group,dtype", [((24, 512), 128, torch.bfloat16), ((7, 1024), 64, torch.bfloat16), ((33, 256), 32, torch.float16),
                                               ((16, 512), 128, torch.float32), ((40, 96), 8, torch.bfloat16), ((5, 64), 4, torch.float32)])
@pytest.mark.parametrize("symmetric,one_sided", [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize("variant", ["mixed", "positive", "some_positive"])
def test_group_tiles_int8_codes_vs_oracle(shape, group, dtype, symmetric, one_sided, variant):
    tile = (1, group)
    assert ops.calibrate_quantize_mode(shape, tile, dtype) == 3
    g = torch.Generator().manual_seed(sum(shape) + group + symmetric * 4 + one_sided * 2 + len(variant))
    xs = _variants(shape, g, variant, dtype, 3)
    outs, flags, _ = _run_fused(xs, tile, 4, symmetric, one_sided, rowsum=False)
    assert flags == 0
    mn = mx = None
    for x, (q, rs, scale, offset, rmn, rmx) in zip(xs, outs):
        mn, mx, s, o, rq = _oracle_step(mn, mx, x, tile, 4, symmetric, one_sided)
        assert rs is None
        assert bits_equal(rmn, mn) and bits_equal(rmx, mx) and bits_equal(scale, s)
        if offset is not None:
            assert bits_equal(offset, o if o is not None else torch.zeros_like(s))
        assert bits_equal(q, rq)


FQ_CASES = [((24, 512), (1, 128), torch.bfloat16), ((300, 4096), (1, 128), torch.bfloat16), ((9, 1024), (1, 1024), torch.bfloat16),
            ((64, 4096), (1, 4096), torch.float32), ((5, 14336), (1, 14336), torch.bfloat16), ((17, 256), (1, 32), torch.float16),
            ((6, 20480), (1, 20480), torch.bfloat16)]


@pytest.mark.parametrize("shape,tile,dtype", FQ_CASES)
@pytest.mark.parametrize("symmetric,one_sided", [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize("variant", ["mixed", "positive", "some_positive"])
@pytest.mark.parametrize("bits,qdtype", [(4, None), (8, torch.int8)])
def test_calibrate_fake_quantize_vs_oracle(shape, tile, dtype, symmetric, one_sided, variant, bits, qdtype):
    """min/max -> params -> dequantize(quantize(x)) in one launch, in place, against the oracle's three steps."""
    g = torch.Generator().manual_seed(sum(shape) + tile[1] + symmetric * 4 + one_sided * 2 + len(variant) + bits)
    x = _variants(shape, g, variant, dtype, 1)[0]
    x.view(-1)[3] = -0.0
    nt = (shape[0] // tile[0]) * (shape[1] // tile[1])
    mn, mx = R.tile_minmax(x, tile)
    s, o = R.parameters_for_range(mn, mx, bits, symmetric, one_sided)
    q = R.quantize_by_tile(x, s, tile, bits, qdtype or x.dtype, o)
    want = R.dequantize_by_tile(q, s, tile, o, x.dtype)
    scale = torch.empty(nt, device=DEV)
    offset = None if (symmetric and not one_sided) else torch.empty(nt, device=DEV)
    w = x.to(DEV)
    before = ff._cabi.launch_count()
    out = ops.calibrate_fake_quantize_(w, tile, bits, symmetric, one_sided, scale, offset, qdtype, out=w)
    assert ff._cabi.launch_count() - before <= 2 and out.data_ptr() == w.data_ptr()
    assert bits_equal(scale, s)
    if offset is not None:
        assert bits_equal(offset, o if o is not None else torch.zeros_like(s))
    assert bits_equal(w, want)


def test_calibrate_fake_quantize_running_range_and_special_values():
    g = torch.Generator().manual_seed(4)
    x1 = torch.randn(12, 1024, generator=g); x2 = torch.randn(12, 1024, generator=g) * 3
    x2[2, 5] = float("nan"); x2[3, :] = 0.0; x2[4, 7] = 1e30
    for tile in [(1, 128), (1, 1024)]:
        nt = 12 * (1024 // tile[1])
        for sym in (True, False):
            mn = torch.full((nt,), float("inf"), device=DEV); mx = -mn
            scale, offset = torch.empty(nt, device=DEV), torch.empty(nt, device=DEV)
            rmn = rmx = None
            for x in (x1, x2):
                out = ops.calibrate_fake_quantize_(x.to(DEV), tile, 8, sym, True, scale, offset, None, run_min=mn, run_max=mx)
                rmn, rmx = R.running_minmax_step(rmn, rmx, x, tile)
                s, o = R.parameters_for_range(rmn, rmx, 8, sym, True)
                want = R.dequantize_by_tile(R.quantize_by_tile(x, s, tile, 8, x.dtype, o), s, tile, o, x.dtype)
                assert bits_equal(mn, rmn) and bits_equal(mx, rmx) and bits_equal(scale, s)
                assert bits_equal(offset, o if o is not None else torch.zeros_like(s))
                assert bits_equal(out, want)


def test_calibrate_and_fuse_qdq_weights_equals_two_step():
    """Whole-model weight quantization (BASELINE config 3): the fused per-weight launch gives the same weights and
    quantizer parameters as calibrate_weight_quantizers + fuse_qdq_weights, with fewer launches."""
    from fastforward_b200.quantization.fuse import calibrate_and_fuse_qdq_weights, calibrate_weight_quantizers, fuse_qdq_weights

    def build(gran, bits):
        torch.manual_seed(1)
        m = torch.nn.Sequential(torch.nn.Linear(512, 256, bias=False), torch.nn.Linear(256, 384, bias=True)).to(DEV).bfloat16()
        with torch.no_grad():
            m[1].weight.abs_()                                   # every tile non-negative: the one-sided branch
        ff.quantize_model(m)
        ff.find_quantizers(m, "**/[quantizer:parameter/weight]").initialize(ff.nn.LinearQuantizer, num_bits=bits, granularity=gran())
        return m.to(DEV)

    for gran, bits in [(lambda: ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0), 4), (lambda: ff.PerChannel(0), 8),
                       (lambda: ff.PerTensor(), 8)]:
        a, b = build(gran, bits), build(gran, bits)
        l0 = ff._cabi.launch_count()
        calibrate_and_fuse_qdq_weights(a)
        l1 = ff._cabi.launch_count()
        calibrate_weight_quantizers(b); fuse_qdq_weights(b)
        l2 = ff._cabi.launch_count()
        for (na, pa), (nb, pb) in zip(a.state_dict().items(), b.state_dict().items()):
            assert na == nb and bits_equal(pa, pb), na
        assert (l1 - l0) <= (l2 - l1)


def test_estimator_group_tiles_one_group_per_row():
    """A per-group quantizer whose group spans the whole (short) row takes the group kernel, which has no row sums."""
    q = ff.nn.LinearQuantizer(4, granularity=ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0),
                              quantized_dtype=torch.int8, device=DEV)
    x = torch.randn(256, 128, device=DEV).bfloat16()
    with torch.no_grad(), ff.estimate_ranges(q, ff.range_setting.running_minmax):
        out = q(x)
    mn, mx, s, o, rq = _oracle_step(None, None, x.cpu(), (1, 128), 4, True, True)
    assert bits_equal(out.raw_data, rq) and bits_equal(q.scale.detach(), s) and not hasattr(out, "_ffq_rowsum")
