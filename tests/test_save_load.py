"""On-disk formats: quantization state and quantized-model artifacts.

CPU tests: (1) files the UNMODIFIED reference wrote (tests/golden/formats, recorded by oracle/make_golden_formats.py)
load into this package with the expected settings and tensors; (2) round trips inside this package incl. shared
quantizers, tied weights, lazy parameters, overwrite policies and the error conventions of the reference's own
tests/quantization/test_save_load.py; (3) wherever the staged reference is present (oracle/_ref), files written by this
package are read back by the reference itself (a subprocess, so the two packages' import hooks never meet)."""
import json
import os
import subprocess
import sys
import textwrap

import pytest
import torch

import fastforward_b200 as ff
from fastforward_b200 import serialization

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FORMATS = os.path.join(ROOT, "tests", "golden", "formats")


def _quantized_toy(seed=11):
    torch.manual_seed(seed)
    model = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.Linear(8, 4))
    ff.quantize_model(model)
    return model


def _initialised_toy(seed=11):
    """Quantizers with materialised parameters, no device needed: the values are set directly."""
    model = _quantized_toy(seed)
    ff.find_quantizers(model, "**/[quantizer:parameter/weight]").initialize(
        ff.nn.LinearQuantizer, num_bits=4, granularity=ff.PerBlock(block_dims=1, block_sizes=4, per_channel_dims=0))
    ff.find_quantizers(model, "**/[quantizer:activation/input]").initialize(
        ff.nn.LinearQuantizer, num_bits=8, symmetric=False, granularity=ff.PerTensor(), quantized_dtype=torch.int8)
    g = torch.Generator().manual_seed(seed)
    for name, q in ff.nn.named_quantizers(model):
        n = 1 if "input" in name else (32 if name.startswith("0.") else 8)
        q._initialize_parameters(n)
        with torch.no_grad():
            q.scale.copy_(torch.rand(n, generator=g) + 0.1)
            if q.offset is not None:
                q.offset.copy_(torch.randint(-3, 4, (n,), generator=g).float())
    return model


def _assert_same_quantizers(a, b):
    qa, qb = dict(ff.nn.named_quantizers(a)), dict(ff.nn.named_quantizers(b))
    assert qa.keys() == qb.keys() and qa
    for name in qa:
        x, y = qa[name], qb[name]
        assert type(x) is type(y) and x.num_bits == y.num_bits and x.symmetric == y.symmetric
        assert x.granularity == y.granularity and x.quantized_dtype == y.quantized_dtype
        assert torch.equal(x.scale, y.scale)
        assert (x.offset is None) == (y.offset is None)
        if x.offset is not None:
            assert torch.equal(x.offset, y.offset)


# ---- (1) the reference's files --------------------------------------------------------------------------------
def test_reference_written_state_loads():
    expect = torch.load(os.path.join(FORMATS, "expect.pt"), weights_only=False)
    import types
    model = _quantized_toy()
    with pytest.raises(RuntimeError, match="model identifier"):      # a path says where, not for which model
        ff.quantization.load_quantization_state(model, name_or_path=os.path.join(FORMATS, "ref_state", "config.yaml"))
    model.config = types.SimpleNamespace(name_or_path="toy/model")
    ff.quantization.load_quantization_state(model, name_or_path=os.path.join(FORMATS, "ref_state", "config.yaml"))
    got = dict(ff.nn.named_quantizers(model))
    assert got.keys() == expect["quantizers"].keys()
    for name, want in expect["quantizers"].items():
        q = got[name]
        assert isinstance(q, ff.nn.LinearQuantizer)
        assert (q.num_bits, q.symmetric, q.allow_one_sided, q.quantized_dtype) == \
            (want["num_bits"], want["symmetric"], want["allow_one_sided"], want["quantized_dtype"])
        assert repr(q.granularity).split("(")[0] == want["granularity"].split("(")[0]
        assert torch.equal(q.scale.detach(), want["scale"])
        assert (q.offset is None) == (want["offset"] is None)
        if want["offset"] is not None:
            assert torch.equal(q.offset.detach(), want["offset"])
        assert q.quant_metadata is not None          # the slot's metadata is re-attached


def test_reference_written_artifact_loads():
    expect = torch.load(os.path.join(FORMATS, "expect.pt"), weights_only=False)
    model = _quantized_toy(seed=99)                   # other weights: the artifact's must replace them
    ff.quantization.load_quantized_model(model, os.path.join(FORMATS, "ref_artifact"), expected_name="toy/model")
    state = model.state_dict()
    assert state.keys() == expect["weights"].keys()
    for k, v in expect["weights"].items():
        assert torch.equal(state[k], v), k
    with pytest.raises(RuntimeError, match="Model identifier mismatch"):
        ff.quantization.load_quantized_model(_quantized_toy(), os.path.join(FORMATS, "ref_artifact"), expected_name="other")


def test_config_yaml_has_the_reference_form(tmp_path):
    """Same tag, same keys, class names spelled with the reference's root package."""
    model = _initialised_toy()
    cfg = ff.quantization.save_quantization_state(model, name_or_path="toy/model", cache_dir=tmp_path)
    assert cfg == tmp_path.resolve() / "quantization-state--toy--model" / "main" / "config.yaml"
    text = cfg.read_text()
    ref_text = open(os.path.join(FORMATS, "ref_state", "config.yaml")).read()
    assert "!ff.obj" in text and "fastforward_b200" not in text
    assert "name: fastforward.nn.linear_quantizer.LinearQuantizer" in text
    assert "name: fastforward.quantization.granularity.PerBlock" in text and "name: torch.int8" in text
    keys = lambda t: {ln.split(":")[0].strip() for ln in t.splitlines() if ln and not ln.lstrip().startswith("-")}  # noqa: E731
    assert {"version", "name_or_path", "transformers_version", "fastforward_version", "quantizers", "name", "initargs",
            "state", "num_bits", "granularity", "quantized_dtype", "allow_one_sided"} <= keys(text) & keys(ref_text)
    from safetensors import safe_open
    with safe_open(str(cfg.parent / "model.safetensors"), framework="pt") as f:
        meta = f.metadata()
        assert meta["0.weight_quantizer"] == "scale=0.weight_quantizer.scale,offset=0.weight_quantizer.offset"
        assert set(f.keys()) == {f"{n}.{p}" for n, _ in ff.nn.named_quantizers(model) for p in ("scale", "offset")}


# ---- (2) round trips --------------------------------------------------------------------------------------------
def test_state_round_trip_and_policies(tmp_path):
    model = _initialised_toy()
    ff.quantization.save_quantization_state(model, tag="t1", name_or_path="m", cache_dir=tmp_path)
    fresh = _quantized_toy()
    ff.quantization.load_quantization_state(fresh, tag="t1", name_or_path="m", cache_dir=tmp_path)
    _assert_same_quantizers(model, fresh)
    # loading again over initialised quantizers: error / skip / overwrite (reference save_load.py:511-537)
    with pytest.raises(ff.QuantizationError, match="already initialized"):
        ff.quantization.load_quantization_state(fresh, tag="t1", name_or_path="m", cache_dir=tmp_path)
    kept = fresh[0].weight_quantizer
    ff.quantization.load_quantization_state(fresh, tag="t1", name_or_path="m", cache_dir=tmp_path, overwrite_policy="skip")
    assert fresh[0].weight_quantizer is kept
    ff.quantization.load_quantization_state(fresh, tag="t1", name_or_path="m", cache_dir=tmp_path, overwrite_policy="overwrite")
    assert fresh[0].weight_quantizer is not kept
    _assert_same_quantizers(model, fresh)
    with pytest.raises(ff.QuantizationError):
        ff.quantization.load_quantization_state(fresh, tag="t1", name_or_path="m", cache_dir=tmp_path, overwrite_policy="bogus")
    with pytest.raises(FileNotFoundError):
        ff.quantization.load_quantization_state(_quantized_toy(), tag="absent", name_or_path="m", cache_dir=tmp_path)
    with pytest.raises(RuntimeError, match="model identifier"):
        ff.quantization.save_quantization_state(model, cache_dir=tmp_path)      # no config.name_or_path, none given


def test_methods_on_quantized_modules_and_model_config_identifier(tmp_path):
    import types
    model = _initialised_toy()
    model.config = types.SimpleNamespace(name_or_path="org/net", transformers_version="9.9")
    cfg = model[0].save_quantization_state(name_or_path="org/net", cache_dir=tmp_path)     # method form
    assert cfg.exists()
    cfg2 = ff.quantization.save_quantization_state(model, cache_dir=tmp_path, tag="v2")   # identifier from model.config
    loaded = serialization.load(open(cfg2))
    assert loaded["name_or_path"] == "org/net" and loaded["transformers_version"] == "9.9" and loaded["version"] == "1.0"
    other = _quantized_toy()
    other.config = types.SimpleNamespace(name_or_path="someone/else")
    with pytest.raises(RuntimeError, match="Model identifier mismatch"):
        ff.quantization.load_quantization_state(other, name_or_path=cfg2)


def test_shared_quantizer_is_stored_once(tmp_path):
    model = _initialised_toy()
    model[1].input_quantizer = model[0].input_quantizer           # one instance, two places
    cfg = ff.quantization.save_quantization_state(model, name_or_path="m", cache_dir=tmp_path)
    from safetensors import safe_open
    with safe_open(str(cfg.parent / "model.safetensors"), framework="pt") as f:
        assert "1.input_quantizer.scale" not in f.keys() and "0.input_quantizer.scale" in f.keys()
        assert f.metadata()["1.input_quantizer"] == "scale=0.input_quantizer.scale,offset=0.input_quantizer.offset"
    fresh = _quantized_toy()
    ff.quantization.load_quantization_state(fresh, name_or_path="m", cache_dir=tmp_path)
    assert fresh[1].input_quantizer is fresh[0].input_quantizer    # YAML anchors keep the sharing
    _assert_same_quantizers(model, fresh)


def test_lazy_parameters(tmp_path):
    model = _quantized_toy()
    ff.find_quantizers(model, "**/[quantizer:parameter/weight]").initialize(ff.nn.LinearQuantizer, num_bits=8)
    with pytest.raises(ValueError, match="lazy parameters"):
        ff.quantization.save_quantization_state(model, name_or_path="m", cache_dir=tmp_path)
    cfg = ff.quantization.save_quantization_state(model, name_or_path="m", cache_dir=tmp_path, allow_lazy_params=True)
    from safetensors import safe_open
    with safe_open(str(cfg.parent / "model.safetensors"), framework="pt") as f:
        assert f.metadata()["0.weight_quantizer"] == "scale=0.weight_quantizer.scale::lazy,offset=0.weight_quantizer.offset::lazy"
    with pytest.raises(ValueError, match="Lazy parameters"):
        ff.quantization.load_quantization_state(_quantized_toy(), name_or_path="m", cache_dir=tmp_path)
    fresh = _quantized_toy()
    ff.quantization.load_quantization_state(fresh, name_or_path="m", cache_dir=tmp_path, allow_lazy_params=True)
    assert fresh[0].weight_quantizer.has_uninitialized_params


def test_model_without_quantizers_is_readable(tmp_path):
    model = _quantized_toy()                                       # stubs only
    ff.quantization.save_quantized_model(model, tmp_path / "a", name_or_path="m")
    fresh = _quantized_toy(seed=5)
    ff.quantization.load_quantized_model(fresh, tmp_path / "a", expected_name="")
    assert torch.equal(fresh[0].weight, model[0].weight)


def test_artifact_round_trip_with_tied_weights(tmp_path):
    model = _initialised_toy()
    tied = torch.nn.Linear(16, 8)
    model.add_module("extra", tied)
    tied.weight = model[0].weight                                  # tied parameters: safetensors needs them split
    path = ff.quantization.save_quantized_model(model, tmp_path / "art", name_or_path="m")
    manifest = json.load(open(path / "manifest.json"))
    assert manifest["version"] == "1.0" and manifest["tied_weights"] == {"extra.weight": "0.weight"}
    assert sorted(os.listdir(path)) == ["config.yaml", "manifest.json", "quantizer_state.safetensors", "weights.safetensors"]
    fresh = _quantized_toy(seed=3)
    fresh.add_module("extra", torch.nn.Linear(16, 8))
    ff.quantization.load_quantized_model(fresh, path, expected_name="m")
    _assert_same_quantizers(model, fresh)
    for k, v in model.state_dict().items():
        assert torch.equal(fresh.state_dict()[k], v), k
    # a model with other weight keys is refused, a missing file is named
    with pytest.raises(RuntimeError, match="do not match this model"):
        ff.quantization.load_quantized_model(_quantized_toy(), path, expected_name="m")
    os.remove(path / "weights.safetensors")
    with pytest.raises(FileNotFoundError, match="weights.safetensors"):
        ff.quantization.load_quantized_model(_quantized_toy(), path)


def test_overrides_block_serialisation(tmp_path):
    model = _initialised_toy()
    with model[0].weight_quantizer.register_override(lambda q, cb, a, k: cb(*a, **k)):
        with pytest.raises(RuntimeError, match="overrides"):
            ff.quantization.save_quantization_state(model, name_or_path="m", cache_dir=tmp_path)


def test_user_defined_quantizer_subclass_round_trips(tmp_path):
    q = _UserQuantizer(5, extra="abc", granularity=ff.PerChannel(1))
    text = serialization.dump({"q": q})
    assert f"name: {__name__}._UserQuantizer" in text
    back = serialization.load(text)["q"]
    assert isinstance(back, _UserQuantizer) and back.extra == "abc" and back.num_bits == 5
    assert back.granularity == ff.PerChannel(1)


class _UserQuantizer(ff.nn.LinearQuantizer):
    def __init__(self, num_bits, extra="x", **kw):
        super().__init__(num_bits, **kw)
        self.extra = extra


# ---- (3) the reference reads what this package wrote --------------------------------------------------------------
def test_reference_reads_our_files(tmp_path):
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("the staged reference (oracle/_ref) is not present")
    model = _initialised_toy()
    ff.quantization.save_quantization_state(model, name_or_path="toy/model", cache_dir=tmp_path / "cache")
    ff.quantization.save_quantized_model(model, tmp_path / "art", name_or_path="toy/model")
    want = tmp_path / "want.pt"
    torch.save({"q": {n: (q.scale.detach(), None if q.offset is None else q.offset.detach(), q.num_bits, q.symmetric)
                      for n, q in ff.nn.named_quantizers(model)},
                "w": {k: v.detach() for k, v in model.state_dict().items()}}, want)
    script = textwrap.dedent(f"""
        import sys, torch
        sys.path.insert(0, {ROOT!r})
        from oracle.ref_loader import load_reference
        ff = load_reference()
        want = torch.load({str(want)!r}, weights_only=False)
        def toy():
            torch.manual_seed(1)
            m = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.Linear(8, 4))
            ff.quantize_model(m)
            return m
        m = toy()
        ff.quantization.load_quantization_state(m, name_or_path="toy/model", cache_dir={str(tmp_path / 'cache')!r})
        got = dict(ff.nn.quantized_module.named_quantizers(m))
        assert got.keys() == want["q"].keys(), (got.keys(), want["q"].keys())
        for n, (s, o, bits, sym) in want["q"].items():
            q = got[n]
            assert type(q).__module__.startswith("fastforward.") and q.num_bits == bits and q.symmetric == sym
            assert torch.equal(q.scale.detach(), s) and (o is None or torch.equal(q.offset.detach(), o))
        m = toy()
        ff.quantization.load_quantized_model(m, {str(tmp_path / 'art')!r}, expected_name="toy/model")
        for k, v in want["w"].items():
            assert torch.equal(m.state_dict()[k], v), k
        print("REFERENCE-READ-OK")
    """)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "REFERENCE-READ-OK" in out.stdout, out.stderr[-3000:]
