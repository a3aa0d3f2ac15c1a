"""Host-side logic of the drop-in layer: no GPU, no compute calls.  Mirrors the reference's own
tests for these pieces (tests/test_dispatcher.py, tests/quantization/test_tiled_tensor.py,
tests/quantization/test_granularity.py, tests/nn/test_quantizer.py, tests/test_flags.py)."""
import ctypes
import os
import re

import pytest
import torch

import fastforward_b200 as ff
from fastforward_b200 import _cabi, dispatcher
from fastforward_b200.quantization import granularity as G
from fastforward_b200.quantization import tiled_tensor as TT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI -----------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ffq_b200.h")).read()
    declared = sorted(set(re.findall(r"FFQ_API\s+[\w\s\*]+?\b(ffq_\w+)\s*\(", header)))
    assert len(declared) >= 15
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/ffq_b200.h but not exported"
    assert sorted(_cabi.EXPORTED) == declared
    assert lib.ffq_abi_version() == _cabi.ABI_VERSION == 2


def test_layout_planning_through_the_abi():
    lay = _cabi.make_layout((32, 16, 8), (1, 16, 8))
    assert _cabi.lib.ffq_num_tiles(ctypes.byref(lay)) == 32
    lay = _cabi.make_layout((4096, 4096), (1, 128))
    assert _cabi.lib.ffq_num_tiles(ctypes.byref(lay)) == 4096 * 32
    assert _cabi.workspace_bytes(_cabi.WS_QUANTIZE_BWD, lay, torch.float32) == 0
    per_tensor = _cabi.make_layout((4096, 4096), (4096, 4096))
    assert _cabi.workspace_bytes(_cabi.WS_QUANTIZE_BWD, per_tensor, torch.float32) > 0
    with pytest.raises(ValueError):
        _cabi.make_layout((4, 6), (3, 3))
    with pytest.raises(ValueError):
        _cabi.make_layout((4, 6), (2,))


def test_no_cpu_fallback():
    x = torch.randn(4, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ff.ops.quantize_by_tile(x, torch.ones(1), (4, 4), 8.0, None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ff.ops.tile_minmax(x, (4, 4))
    q = ff.nn.LinearQuantizer(8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        q.quantization_range = (-1.0, 1.0)


# ---- tiles / granularity ------------------------------------------------------------------------
def test_tile_row_permutation():
    data = torch.arange(24).reshape(4, 6)
    rows = TT.tiles_to_rows(data, (2, 3))
    assert rows.tolist() == [[0, 1, 2, 6, 7, 8], [3, 4, 5, 9, 10, 11], [12, 13, 14, 18, 19, 20], [15, 16, 17, 21, 22, 23]]
    assert torch.equal(TT.rows_to_tiles(rows, (4, 6), (2, 3)), data)
    assert torch.equal(TT.tiles_to_rows(data, "data_shape"), data.reshape(1, -1))
    for bad in ((3, 3), (2,)):
        with pytest.raises(ValueError):
            TT.tiles_to_rows(data, bad)
    with pytest.raises(ValueError):
        TT.rows_to_tiles(rows[:2], (4, 6), (2, 3))


def test_granularities():
    shape = torch.Size((32, 16, 8))
    assert G.PerTensor().tile_size(shape) == "data_shape" and G.PerTensor().parameter_dimensionality(shape) == 1
    assert G.PerChannel(0).tile_size(shape) == (1, 16, 8) and G.PerChannel(0).parameter_dimensionality(shape) == 32
    assert G.PerChannel((0, 2)).tile_size(shape) == (1, 16, 1)
    pb = G.PerBlock(block_dims=1, block_sizes=8, per_channel_dims=0)
    assert pb.tile_size(shape) == (1, 8, 8) and pb.parameter_dimensionality(shape) == 64
    with pytest.raises(ValueError):
        G.PerBlock(1, 5, 0).tile_size(shape)
    with pytest.raises(ValueError):
        G.PerBlock(1, 32, 0).tile_size(shape)
    assert G.PerTile((16, 8, 4)).tile_size(shape) == (16, 8, 4)
    with pytest.raises(ValueError):
        G.PerTile((5, 8, 4)).tile_size(shape)
    assert G.PerChannel(1) == G.PerChannel(1) and G.PerChannel(1) != G.PerChannel(0) and G.PerTensor() == G.PerTensor()
    assert G.granularity_from_sizes(shape, shape) == G.PerTensor()
    assert G.granularity_from_sizes(shape, torch.Size((1, 16, 8))) == G.PerChannel(0)
    assert G.granularity_from_sizes(shape, torch.Size((1, 8, 8))).tile_size(shape) == (1, 8, 8)


# ---- dispatcher ----------------------------------------------------------------------------------
@pytest.fixture
def clean_registry():
    saved = {k: list(v) for k, v in dispatcher._DISPATCHER.items()}
    yield
    dispatcher._DISPATCHER.clear()
    dispatcher._DISPATCHER.update(saved)


def test_dispatcher_order_and_scoping(clean_registry):
    k1, k2, k3 = (lambda *a, **k: 1), (lambda *a, **k: 2), (lambda *a, **k: 3)
    P = ff.Predicate
    dispatcher.register("my_op", P(lambda x: x > 0), k1)
    dispatcher.register("my_op", P(lambda x: x > 10), k2)
    assert dispatcher.dispatch("my_op", 5) is k1
    assert dispatcher.dispatch("my_op", 50) is k2          # newest first within a priority
    assert dispatcher.dispatch("my_op", -1) is None
    with dispatcher.register("my_op", None, k3):
        assert dispatcher.dispatch("my_op", -1) is k3
    assert dispatcher.dispatch("my_op", -1) is None        # scoped registration removed
    dispatcher.register("my_op", None, k3, dispatcher.DispatcherPriority.FALLBACK)
    assert dispatcher.dispatch("my_op", 50) is k2 and dispatcher.dispatch("my_op", -1) is k3

    @dispatcher.register("other", P(lambda **kw: kw.get("flag", False)))
    def kern(**kw):
        return "k"

    assert dispatcher.dispatch("other", flag=True) is kern and dispatcher.dispatch("other", flag=False) is None
    both = P(lambda x: x > 0) & ~P(lambda x: x > 10)
    assert both(5) and not both(50) and (P(lambda x: x < 0) | both)(-3)


def test_linear_consults_dispatcher_then_fallback(clean_registry):
    calls = []

    def kernel(input, weight, bias, output_quantizer, strict_quantization):
        calls.append("kernel")
        return torch.zeros(1)

    x, w = torch.randn(2, 4), torch.randn(3, 4)
    with dispatcher.register("linear", ff.Predicate(lambda **kw: kw["bias"] is None), kernel):
        ff.nn.functional.linear(x, w, None, strict_quantization=False)
        y = ff.nn.functional.linear(x, w, torch.zeros(3), strict_quantization=False)   # predicate rejects
    assert calls == ["kernel"] and torch.equal(y, torch.nn.functional.linear(x, w))
    with pytest.raises(ff.QuantizationError):
        ff.nn.functional.linear(x, w, None, strict_quantization=True)   # strict: needs output_quantizer


# ---- flags -----------------------------------------------------------------------------------------
def test_flags():
    assert ff.get_strict_quantization() is True and ff.get_export_mode() is False
    with ff.strict_quantization(False):
        assert ff.get_strict_quantization() is False
    assert ff.get_strict_quantization() is True
    restore = ff.set_export_mode(True)
    assert ff.get_export_mode() is True
    with restore:
        pass
    assert ff.get_export_mode() is False


# ---- quantizers, overrides, model conversion ----------------------------------------------------------
def test_override_stack_runs_newest_outermost():
    q = ff.nn.QuantizerStub()
    order = []

    def make(tag):
        def ov(quantizer, nxt, args, kwargs):
            order.append(tag)
            return nxt(*args, **kwargs)
        return ov

    h1, h2 = q.register_override(make("first")), q.register_override(make("second"))
    x = torch.ones(2)
    assert q(x) is x and order == ["second", "first"]
    h2.remove()
    order.clear()
    q(x)
    assert order == ["first"]
    with q.register_override(make("scoped")):
        order.clear()
        q(x)
        assert order == ["scoped", "first"]
    h1.remove()
    assert list(q.overrides) == []


def test_tags_and_metadata():
    T = ff.nn.Tag
    assert T("a/b") is T("a/b") and str(T("a") / "b") == "#a/b"
    meta = ff.nn.QuantizerMetadata(weight_quantizer=True, shape=(3, 4))
    assert meta.weight_quantizer and meta.parameter_quantizer and not meta.input_quantizer
    assert "parameter/weight" in meta and "parameter" in meta and meta.shape == (3, 4)
    assert ff.nn.QuantizerMetadata(weight_quantizer=True).is_extension(meta)


def test_linear_quantizer_lazy_parameters():
    q = ff.nn.LinearQuantizer(8, granularity=ff.PerChannel())
    assert q.symmetric and isinstance(q.scale, torch.nn.UninitializedParameter)
    assert isinstance(q.offset, torch.nn.UninitializedBuffer) and q.quantization_range == (None, None)
    assert ff.nn.LinearQuantizer(8, allow_one_sided=False).offset is None
    asym = ff.nn.LinearQuantizer(4, symmetric=False)
    assert isinstance(asym.offset, torch.nn.UninitializedParameter) and not asym.symmetric
    with pytest.raises(ValueError, match="uninitialized quantizer"):
        q(torch.randn(3, 3))
    assert q.integer_minimum == -128 and q.integer_maximum == 127 and asym.integer_maximum == 7
    q._initialize_parameters(5)
    assert q.scale.shape == (5,) and torch.all(q.scale == 1) and torch.all(q.offset == 0)
    q.reset_parameters()
    assert q.has_uninitialized_params
    sd = {"scale": torch.full((7,), 0.5), "offset": torch.zeros(7)}
    q.load_state_dict(sd)                       # lazy params take the loaded shape
    assert q.scale.shape == (7,) and float(q.scale[0]) == 0.5


def test_quantize_model_and_find_quantizers():
    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2 = torch.nn.Linear(8, 16), torch.nn.Linear(16, 8, bias=False)
            self.act = torch.nn.SiLU()

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    model = torch.nn.Sequential(Block(), Block())
    with pytest.raises(ff.QuantizationError):
        ff.quantize_model(model)                 # Block has no quantized counterpart
    extra = ff.surrogate_quantized_modules(model)
    ff.quantize_model(model, extra_conversion=extra)
    assert isinstance(model[0].fc1, ff.nn.QuantizedLinear) and type(model[0]).__name__ == "QuantizedBlockSurrogate"
    assert model[0].fc2.bias_quantizer is None
    assert list(ff.nn.named_quantizers(model)) == []          # stubs are skipped
    weights = ff.find_quantizers(model, "**/[quantizer:parameter/weight]")
    assert sorted(r.full_name for r in weights) == ["0.fc1.weight_quantizer", "0.fc2.weight_quantizer",
                                                    "1.fc1.weight_quantizer", "1.fc2.weight_quantizer"]
    weights.initialize(ff.nn.LinearQuantizer, num_bits=8, granularity=ff.PerChannel())
    assert isinstance(model[1].fc2.weight_quantizer, ff.nn.LinearQuantizer)
    assert model[1].fc2.weight_quantizer.quant_metadata.weight_quantizer      # slot metadata re-attached
    with pytest.raises(ff.QuantizationError):
        ff.find_quantizers(model, "**/[quantizer:parameter/weight]").initialize(ff.nn.LinearQuantizer, num_bits=4)
    ff.find_quantizers(model, "**/[quantizer:parameter/weight]").initialize(
        ff.nn.LinearQuantizer, num_bits=4, overwrite_policy="skip")
    assert model[0].fc1.weight_quantizer.num_bits == 8
    ff.find_quantizers(model, "0/fc1/[quantizer:activation/input]").initialize(
        lambda name, cur: ff.nn.LinearQuantizer(8, symmetric=False))
    assert len(list(ff.nn.named_quantizers(model))) == 5
    assert len(ff.find_quantizers(model, "*/[cls:torch.nn.Linear]/[quantizer:activation]")) == 8
    assert len(ff.find_quantizers(model, "**/[re:fc\\d]/weight_quantizer")) == 4
    assert len(ff.find_quantizers(model, "**/~[quantizer:parameter]")) == 10

    cfg = ff.QuantizationConfig()
    cfg.add_rule("**/[quantizer:activation/output]", ff.nn.LinearQuantizer, num_bits=8)
    cfg.add_rule("1/**/[quantizer:activation/output]", ff.nn.LinearQuantizer, num_bits=4)   # last rule wins
    cfg.initialize(model)
    assert model[0].fc1.output_quantizer.num_bits == 8 and model[1].fc1.output_quantizer.num_bits == 4


def test_disable_quantization_restores_flags_and_overrides():
    model = torch.nn.Sequential(torch.nn.Linear(4, 4))
    ff.quantize_model(model)
    ff.find_quantizers(model, "**/[quantizer:parameter/weight]").initialize(ff.nn.LinearQuantizer, num_bits=8)
    x = torch.randn(2, 4)
    with ff.disable_quantization(model):
        assert ff.get_strict_quantization() is False
        y = model(x)                              # quantizers bypassed: plain float linear, no CUDA needed
    assert torch.allclose(y, torch.nn.functional.linear(x, model[0].weight, model[0].bias))
    assert ff.get_strict_quantization() is True
    assert list(model[0].weight_quantizer.overrides) == []


# ---- host-only pieces of the fused calibration / GPTQ paths (no compute calls) -----------------
def test_calibrate_quantize_mode_classification():
    """Which fused kernel a layout gets is decided on the host from the collapsed tile plan (csrc/ffq_calibrate.cu:
    calq_mode): 1 = a row per CTA, 2 = whole tensor (grid barrier), 3 = sub-warp groups, 0 = the separate kernels."""
    from fastforward_b200 import ops
    bf, f32 = torch.bfloat16, torch.float32
    assert ops.calibrate_quantize_mode((4096, 4096), (1, 4096), bf) == 1           # per-channel weight
    assert ops.calibrate_quantize_mode((4096, 14336), (1, 14336), bf) == 1
    assert ops.calibrate_quantize_mode((2, 3, 4096), (1, 1, 4096), bf) == 1        # leading dims collapse into rows
    assert ops.calibrate_quantize_mode((1, 2048, 4096), "data_shape", bf) == 2     # per-tensor activation
    assert ops.calibrate_quantize_mode((14336, 4096), (1, 128), bf) == 3           # g = 128: 16 vectors
    assert ops.calibrate_quantize_mode((14336, 4096), (1, 128), f32) == 3          # 32 vectors
    assert ops.calibrate_quantize_mode((64, 256), (1, 256), f32) == 1              # 64 vectors: a row again
    assert ops.calibrate_quantize_mode((64, 96), (1, 96), bf) == 0                 # 12 vectors: not a power of two
    assert ops.calibrate_quantize_mode((64, 4096), (64, 1), bf) == 0               # per-channel on the last dim: strided
    assert ops.calibrate_quantize_mode((64, 4096), (8, 128), bf) == 0              # 2-D tiles
    assert ops.calibrate_quantize_mode((8, 40000 * 8), (1, 40000 * 8), bf) == 0    # rows longer than 4096 vectors
    assert ops.calibrate_quantize_mode((64, 4096), (1, 4096), torch.int32) == 0
    assert ops.calibrate_quantize_mode((0, 4096), (1, 4096), bf) == 0
    assert ops.calibrate_quantize_mode((8, 100), (1, 100), bf) == 0                # tile not a multiple of the vector width


def test_params_for_ranges_encode_words():
    words = (ctypes.c_int64 * 3)()
    import struct
    for bits, sym, one_sided in [(8, True, True), (4, False, True), (2, True, False), (16, False, False)]:
        _cabi.lib.ffq_params_for_ranges_encode(float(bits), int(sym), int(one_sided), words)
        lo = -2.0 ** (bits - 1)
        f = struct.unpack("<4f", struct.pack("<2q", words[0], words[1]))
        assert f == (abs(lo), abs(-lo - 1), -lo, 2.0 ** bits - 1)
        assert words[2] == (1 if sym else 0) | (2 if one_sided else 0)


def test_gptq_argument_checks_need_no_gpu():
    from fastforward_b200.quantization import gptq as Gq
    layer = ff.nn.QuantizedLinear(64, 8, bias=False)
    with pytest.raises(ValueError, match="LinearQuantizer"):            # stub weight quantizer (gptq.py:47-49)
        Gq.gptq(layer, [])
    layer.weight_quantizer = ff.nn.LinearQuantizer(4, granularity=G.PerBlock(block_dims=1, block_sizes=16, per_channel_dims=0,
                                                                            strict_blocks=False))
    with pytest.raises(ValueError, match="strict_blocks"):             # gptq.py:59-61
        Gq.gptq(layer, [])
    layer.weight_quantizer = ff.nn.LinearQuantizer(4, granularity=G.PerChannel(0))
    with pytest.raises(NotImplementedError, match="block_size"):
        Gq.gptq(layer, [], block_size=129)
    with pytest.raises(RuntimeError, match="no CPU fallback"):          # a CPU layer never reaches a kernel
        Gq.gptq(layer, [])
    with pytest.raises(TypeError, match="Unsupported granularity"):     # gptq.py:219-221
        Gq._check_granularity(G.PerChannel(2))


# ---- estimate_ranges: block exit protocol (range_setting/common.py) ----------------------------------------------
class _RecordingEstimator:
    """Minimal RangeEstimator: records the order of the calls estimate_ranges makes when the block ends."""

    def __init__(self, runs_cleanup: bool, fail_in_finalize: bool = False):
        self.log = []
        self.finalize_runs_cleanup = runs_cleanup
        self.fail = fail_in_finalize

    def split_module(self, module):
        yield from (m for m in module.modules() if isinstance(m, torch.nn.Linear))

    def prepare(self, module):
        self.log.append(("prepare", id(module)))
        return id(module)

    def cleanup(self, module, metadata):
        assert metadata == id(module)
        self.log.append(("cleanup", id(module)))

    def finalize(self, prepared, cleanup=None):
        self.log.append(("finalize-begin", len(prepared)))
        if cleanup is not None:
            cleanup()                      # host work that needs nothing from the device goes in front of the sync
        if self.fail:
            raise RuntimeError("boom")
        self.log.append(("finalize-end", None))


@pytest.mark.parametrize("runs_cleanup", [True, False])
def test_estimate_ranges_exit_protocol(runs_cleanup):
    model = torch.nn.Sequential(torch.nn.Linear(2, 2), torch.nn.Linear(2, 2))
    est = _RecordingEstimator(runs_cleanup)
    with ff.estimate_ranges(model, est):
        pass
    kinds = [k for k, _ in est.log]
    assert kinds.count("cleanup") == 2 and kinds.count("prepare") == 2            # every module cleaned exactly once
    if runs_cleanup:                                                              # ... inside finalize, before it returns
        assert kinds.index("finalize-end") > max(i for i, k in enumerate(kinds) if k == "cleanup")
    else:                                                                         # ... after finalize (the reference's order)
        assert kinds.index("finalize-end") < min(i for i, k in enumerate(kinds) if k == "cleanup")


def test_estimate_ranges_cleans_up_when_the_block_or_the_exit_fails():
    model = torch.nn.Sequential(torch.nn.Linear(2, 2))
    est = _RecordingEstimator(True)
    with pytest.raises(ValueError):
        with ff.estimate_ranges(model, est):
            raise ValueError("user code failed")
    kinds = [k for k, _ in est.log]
    assert "finalize-begin" not in kinds and kinds.count("cleanup") == 1          # a failed block is not finalized
    for runs_cleanup in (True, False):
        est = _RecordingEstimator(runs_cleanup, fail_in_finalize=True)
        with pytest.raises(RuntimeError, match="boom"):
            with ff.estimate_ranges(model, est):
                pass
        assert [k for k, _ in est.log].count("cleanup") == 1                      # once, whoever ran it


def test_estimate_ranges_argument_contract():
    model = torch.nn.Linear(2, 2)
    with pytest.raises(ValueError, match="already initialized"):
        with ff.estimate_ranges(model, _RecordingEstimator(True), 1):
            pass
    with ff.estimate_ranges([model, model], _RecordingEstimator, True) :        # a class + its arguments, a list of modules
        pass


# ---- overlapped parameter steps: the scheduling logic, with the device replaced by recorders -----------------------
def test_overlapped_parameter_steps_schedule(monkeypatch):
    from fastforward_b200 import ops
    from fastforward_b200.range_setting import minmax as M

    class _Event:
        def record(self, stream=None):
            pass

    class _Stream:
        cuda_stream = 0
        waited = 0

        def wait_event(self, event):
            type(self).waited += 1

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: _Stream())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _Stream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: _Stream())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    where = []

    def fake_calibrate(run_min, run_max, data, tile, *a, stream=None, **k):
        where.append("side" if stream is not None else "main")
        return torch.zeros(data.shape, dtype=torch.int8), None

    monkeypatch.setattr(ops, "calibrate_quantize_", fake_calibrate)
    est = M.RunningMinMaxRangeEstimator(memoize_parameters=False, overlap_parameters=2)
    weights = [torch.nn.Parameter(torch.randn(8, 16)) for _ in range(5)]
    quantizers = [ff.nn.LinearQuantizer(8, granularity=ff.PerChannel(0), quantized_dtype=torch.int8) for _ in weights]
    steps = [M.RunningMinMaxEstimator(q, state=est._state) for q in quantizers]
    with torch.no_grad():
        for s, q, w in zip(steps, quantizers, weights):          # step 1 records the order: everything in line
            s._fused_step(q, w, 1)
        assert where == ["main"] * 5 and est._state.stats["overlapped"] == 0
        where.clear()
        for s, q, w in zip(steps, quantizers, weights):          # step 2: the first in line, the rest launched ahead
            s._fused_step(q, w, 1)
        assert where == ["main", "side", "side", "side", "side"] and est._state.stats["overlapped"] == 4
        assert all(s._ahead is None for s in steps)              # nothing left in flight at the end of a step
        # a weight that changes after its step was launched ahead: the result is dropped, the step re-runs in line
        where.clear()
        steps[0]._fused_step(quantizers[0], weights[0], 1)       # launches 1 and 2 ahead
        assert steps[1]._ahead is not None and steps[2]._ahead is not None
        weights[2].mul_(2.0)                                     # bumps the version counter
        steps[1]._fused_step(quantizers[1], weights[1], 1)
        steps[2]._fused_step(quantizers[2], weights[2], 1)
        assert where == ["main", "side", "side", "side", "main", "side"]     # 0 | 1, 2 ahead | 3 ahead | 2 again | 4 ahead
        # the block ends with launches in flight: drain joins them
        assert steps[3]._ahead is not None and steps[4]._ahead is not None
        est._state.drain()
        assert all(s._ahead is None for s in steps)
    # memoisation makes overlap pointless: it is switched off
    assert M._BlockState(True, True, overlap=4).overlap == 0


# ---- the rest of the `quantization` namespace -----------------------------------------------------------------------
def test_create_quantization_function():
    """reference tests/quantization/test_function.py:116-152."""
    import dataclasses

    data = torch.randn(3, 3)

    def _quantize(data: torch.Tensor, scale: float) -> torch.Tensor:
        return data * scale

    def _dequantize(data: torch.Tensor, rescale: float = 3.5) -> torch.Tensor:
        return data * rescale

    Params, Function, custom = ff.quantization.create_quantization_function("CustomQuantizer", _quantize, _dequantize)
    assert {f.name for f in dataclasses.fields(Params)} == {"scale", "rescale"}
    with ff.strict_quantization(False):
        quantized = custom(data, scale=2.0)
        assert torch.equal(quantized.raw_data, data * 2.0)
        assert torch.equal(quantized.dequantize(), data * 2.0 * 3.5)
    assert quantized.quant_func is Function and isinstance(quantized.quant_args(), Params)
    assert quantized.quant_args().scale == 2.0 and quantized.quant_args().rescale == 3.5

    def _bad_default(data: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        return data

    def _bad_annotation(data: torch.Tensor, scale: int) -> torch.Tensor:
        return data

    def _varargs(data: torch.Tensor, *scales: float) -> torch.Tensor:
        return data

    with pytest.raises(ValueError, match="default value"):
        ff.quantization.create_quantization_function("X", _quantize, _bad_default)
    with pytest.raises(TypeError, match="type annotation"):
        ff.quantization.create_quantization_function("X", _quantize, _bad_annotation)
    with pytest.raises(TypeError, match="keyword"):
        ff.quantization.create_quantization_function("X", _varargs, _dequantize)


def test_quantization_namespace_matches_the_reference():
    """Every public name of the reference's `fastforward.quantization` and the two top-level module aliases."""
    for name in ("dynamic", "static", "freeze_parameters", "create_quantization_function", "ConventionDiscovery",
                 "WeightQuantizerDiscovery", "find_weight_quantizers", "fuse_qdq_weights", "stub_weight_quantizers", "gptq",
                 "QuantizationConfig", "QuantizerCollection", "load_quantization_state", "load_quantized_model",
                 "save_quantization_state", "save_quantized_model"):
        assert hasattr(ff.quantization, name), name
    assert ff.affine is ff.quantization.affine and ff.granularity is ff.quantization.granularity


def test_find_and_stub_weight_quantizers():
    model = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.Linear(4, 2))
    ff.quantize_model(model)
    ff.find_quantizers(model, "**/[quantizer:parameter/weight]").initialize(ff.nn.LinearQuantizer, num_bits=4)
    ff.find_quantizers(model, "**/[quantizer:activation/input]").initialize(ff.nn.LinearQuantizer, num_bits=8)
    targets = ff.quantization.find_weight_quantizers(model)
    assert [(m is model[i], attr) for i, (m, attr, _) in enumerate(targets)] == [(True, "weight"), (True, "weight")]
    assert all(q is m.weight_quantizer for m, _, q in targets)
    assert isinstance(ff.quantization.ConventionDiscovery(), ff.quantization.WeightQuantizerDiscovery)
    before = [m.weight.detach().clone() for m in model]
    meta = model[0].weight_quantizer.quant_metadata
    ff.quantization.stub_weight_quantizers(model)
    assert all(m.weight_quantizer.is_stub() and not m.input_quantizer.is_stub() for m in model)
    assert model[0].weight_quantizer.quant_metadata is meta                  # the slot's metadata survives
    assert all(torch.equal(m.weight, w) for m, w in zip(model, before))      # weights untouched
    assert ff.quantization.find_weight_quantizers(model) == []


def test_quantized_tensor_declines_in_place_operators():
    """reference quantized_tensor.py:488-507: `qt += x` must not write into the codes."""
    from fastforward_b200.quantization.function import QuantizationContext, QuantizationParameters
    qt = ff.QuantizedTensor(torch.zeros(2, 2), QuantizationContext(object, QuantizationParameters()))
    for op in ("__iadd__", "__isub__", "__imul__", "__imatmul__", "__itruediv__", "__ifloordiv__", "__imod__", "__ilshift__",
               "__irshift__", "__iand__", "__ixor__", "__ior__", "__ipow__"):
        assert getattr(qt, op)(1) is NotImplemented, op
