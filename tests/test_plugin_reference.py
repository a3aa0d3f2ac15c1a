"""Drop-in registration against the real, unmodified reference (only where /root/reference exists:
the build container).  No GPU here, so this checks the registration itself: the CUDA-key kernels
appear next to the reference's CompositeExplicitAutograd implementations, the dispatcher receives
the linear kernel, and CPU tensors keep taking the reference's own path."""
import os
import sys
import types

import pytest
import torch

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def ref_ff():
    import torch.utils._pytree as _pt

    optree = types.ModuleType("optree")
    optree.tree_map = lambda fn, tree, *rest, **kw: _pt.tree_map(fn, tree)
    sys.modules.setdefault("optree", optree)
    for name, attrs in (("fastforward.autoquant", {"autoquantize": lambda *a, **k: None}),
                        ("fastforward.testing.autoquant", {})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    sys.path.insert(0, REF)
    import fastforward as ff

    yield ff
    sys.path.remove(REF)


def test_install_registers_cuda_kernels_and_linear(ref_ff):
    from fastforward_b200 import plugin

    before = list(ref_ff.dispatcher._DISPATCHER.get("linear", []))
    installed = plugin.install(ref_ff)
    assert {"quantize_by_tile", "dequantize_by_tile", "quantize_by_tile_backward", "quantize_dynamic_by_tile",
            "linear"} <= set(installed)
    for op in ("quantize_by_tile", "dequantize_by_tile", "quantize_by_tile_backward", "quantize_dynamic_by_tile"):
        dump = torch._C._dispatch_dump(f"fastforward::{op}")
        assert "CUDA" in dump and "CompositeExplicitAutograd" in dump, dump
    assert len(ref_ff.dispatcher._DISPATCHER["linear"]) == len(before) + 1
    # CPU tensors still run the reference's eager chain, bit-identical to the oracle
    from oracle import ref_ops as R

    x = torch.randn(8, 16)
    q = ref_ff.nn.LinearQuantizer(8, granularity=ref_ff.PerChannel(0))
    q.quantization_range = (x.min(1).values, x.max(1).values)
    out = q(x)
    want = R.quantize_by_tile(x, q.scale.detach(), (1, 16), 8, x.dtype, q.offset)
    assert torch.equal(out.raw_data, want)
    # the linear predicate rejects what the kernel does not implement (float codes on CPU): fallback is used
    lin = torch.nn.Linear(16, 4)
    ref_ff.quantize_model(lin)
    lin.input_quantizer = ref_ff.nn.LinearQuantizer(8, symmetric=False)
    lin.weight_quantizer = ref_ff.nn.LinearQuantizer(8, granularity=ref_ff.PerChannel(0))
    lin.input_quantizer.quantization_range = (x.min(), x.max())
    lin.weight_quantizer.quantization_range = (lin.weight.min(1).values, lin.weight.max(1).values)
    with ref_ff.strict_quantization(False):
        y = lin(x)
    assert y.shape == (8, 4)
