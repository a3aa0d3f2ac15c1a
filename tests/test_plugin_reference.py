"""Drop-in registration against the real, unmodified reference staged under oracle/_ref
(``python oracle/fetch_ref.py``; shipped to the GPU box by gpurun).  Without a GPU this checks the registration
itself: the CUDA-key kernels appear next to the reference's CompositeExplicitAutograd implementations, the
reference's dispatcher receives the linear / matmul kernels, and CPU tensors keep taking the reference's own path.
With a GPU (``-m gpu``) the reference's own public API runs on our kernels and is compared with its CPU results."""
import copy

import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not staged (python oracle/fetch_ref.py)")


@pytest.fixture(scope="module")
def ref_ff():
    return ref_loader.load_reference()


def test_install_registers_cuda_kernels_and_linear(ref_ff):
    from fastforward_b200 import plugin

    before = {op: len(ref_ff.dispatcher._DISPATCHER.get(op, [])) for op in ("linear", "matmul", "mm", "bmm")}
    installed = plugin.install(ref_ff)
    assert {"quantize_by_tile", "dequantize_by_tile", "quantize_by_tile_backward", "quantize_dynamic_by_tile",
            "linear"} <= set(installed)
    assert plugin.install(ref_ff) is installed          # idempotent
    for op in ("quantize_by_tile", "dequantize_by_tile", "quantize_by_tile_backward", "quantize_dynamic_by_tile"):
        dump = torch._C._dispatch_dump(f"fastforward::{op}")
        assert "CUDA" in dump and "CompositeExplicitAutograd" in dump, dump
    after = {op: len(ref_ff.dispatcher._DISPATCHER.get(op, [])) for op in before}
    assert after["linear"] == before["linear"] + 2       # W8A8 + W4A16
    assert all(after[op] == before[op] + 1 for op in ("matmul", "mm", "bmm"))
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


def test_estimator_swap_is_reversible(ref_ff):
    from fastforward_b200 import plugin
    from fastforward_b200.range_setting import minmax as ours

    stock = ref_ff.range_setting.running_minmax
    plugin.install_estimators(ref_ff)
    assert ref_ff.range_setting.running_minmax is ours.RunningMinMaxRangeEstimator
    plugin.uninstall_estimators(ref_ff)
    assert ref_ff.range_setting.running_minmax is stock


# ---------------------------------------------------------------------------------------------------------------
# GPU: the reference's public API on our kernels vs the same API on the CPU
# ---------------------------------------------------------------------------------------------------------------
def _w8a8_linear(ff, device, dtype=torch.float32, seed=0):
    torch.manual_seed(seed)
    lin = torch.nn.Linear(256, 96, dtype=dtype)
    ff.quantize_model(lin)
    lin.input_quantizer = ff.nn.LinearQuantizer(8, symmetric=False, quantized_dtype=torch.int8)
    lin.weight_quantizer = ff.nn.LinearQuantizer(8, granularity=ff.PerChannel(0), quantized_dtype=torch.int8)
    return lin.to(device)


@pytest.mark.gpu
def test_reference_api_on_b200_kernels_matches_reference_cpu(ref_ff):
    """quantizer(x) / .dequantize() / backward through the unmodified reference with plugin.install(): bit-exact
    codes, values and dx against the reference's own CPU run; per-tile sums within the reference's test tolerance."""
    from fastforward_b200 import _cabi, plugin

    ff = ref_ff
    plugin.install(ff)
    torch.manual_seed(1)
    for gran, shape in ((ff.PerChannel(0), (64, 512)), (ff.PerTensor(), (32, 256)),
                        (ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0), (16, 512))):
        for symmetric in (True, False):
            x = torch.randn(shape) * 3
            g = torch.randn(shape)
            res = {}
            # the parameters are derived once, on the CPU, and copied: the reference's own range setter gives scales one
            # ulp apart on the two devices (aten's CUDA kernels divide by a scalar as x * (1/d)), which is not under test
            q_cpu = ff.nn.LinearQuantizer(4, symmetric=symmetric, granularity=gran)
            tile = gran.tile_size(x.shape)
            tile = x.shape if isinstance(tile, str) else tile
            rows = ff.quantization.tiled_tensor.tiles_to_rows(x, tile)
            q_cpu.quantization_range = (rows.min(1).values * 0.7, rows.max(1).values * 0.7)
            for device in ("cpu", "cuda"):
                q = copy.deepcopy(q_cpu).to(device)
                xd = x.detach().clone().to(device).requires_grad_()
                l0 = _cabi.launch_count()
                out = q(xd)
                deq = out.dequantize()
                deq.backward(g.to(device))
                res[device] = (out.raw_data.detach().cpu(), deq.detach().cpu(), xd.grad.cpu(), q.scale.grad.cpu(),
                               None if q.offset is None or q.offset.grad is None else q.offset.grad.cpu())
                if device == "cuda":
                    assert _cabi.launch_count() - l0 >= 3, "the CUDA tensors did not reach the ffq kernels"
            c, d = res["cpu"], res["cuda"]
            assert torch.equal(c[0], d[0]) and torch.equal(c[1], d[1]) and torch.equal(c[2], d[2])
            torch.testing.assert_close(d[3], c[3], rtol=1.3e-5, atol=1e-4)
            if c[4] is not None:
                torch.testing.assert_close(d[4], c[4], rtol=1.3e-5, atol=1e-4)


@pytest.mark.gpu
def test_reference_estimate_ranges_and_w8a8_linear_on_b200(ref_ff):
    """ff.estimate_ranges(model, running_minmax) + QuantizedLinear through the unmodified reference: with the stock
    estimator and with the sync-free one, the ranges equal the reference's CPU ranges bit for bit and the W8A8
    tensor-core linear is the kernel the reference's dispatcher selects."""
    from fastforward_b200 import plugin
    from fastforward_b200.nn import qlinear

    ff = ref_ff
    plugin.install(ff)
    batches = [torch.randn(4, 24, 256, generator=torch.Generator().manual_seed(s)) for s in range(3)]
    want = None
    for mode in ("cpu", "cuda-stock-estimator", "cuda-fused-estimator"):
        device = "cpu" if mode == "cpu" else "cuda"
        lin = _w8a8_linear(ff, device)
        if mode == "cuda-fused-estimator":
            plugin.install_estimators(ff)
        calls0 = qlinear.stats()["calls"]
        try:
            with torch.no_grad(), ff.strict_quantization(False), ff.estimate_ranges(lin, ff.range_setting.running_minmax):
                ys = [lin(b.to(device)) for b in batches]
        finally:
            plugin.uninstall_estimators(ff)
        got = dict(x_range=[t.detach().cpu() for t in lin.input_quantizer.quantization_range],
                   w_range=[t.detach().cpu() for t in lin.weight_quantizer.quantization_range],
                   x_scale=lin.input_quantizer.scale.detach().cpu(), w_scale=lin.weight_quantizer.scale.detach().cpu(),
                   y=ys[-1].float().cpu())
        if mode == "cpu":
            want = got
            continue
        assert qlinear.stats()["calls"] - calls0 == len(batches), f"{mode}: the W8A8 kernel was not dispatched"
        for key in ("x_range", "w_range"):
            # min / max are exact on any device; the range getter recomputes them from (scale, offset)
            assert all(torch.allclose(a, b, rtol=1e-6, atol=1e-7) for a, b in zip(got[key], want[key])), f"{mode}: {key}"
        for key in ("x_scale", "w_scale"):
            # aten's CUDA kernels divide by the integer-grid constants as x * (1/d): one ulp from the CPU's x / d
            torch.testing.assert_close(got[key], want[key], rtol=2.5e-7, atol=0)
        # int32-exact accumulation vs the reference's fp32 dequantize-then-GEMM fallback
        torch.testing.assert_close(got["y"], want["y"], rtol=1e-3, atol=1e-3)
        if mode == "cuda-stock-estimator":
            stock = got
        else:
            # the sync-free fused estimator must reproduce the reference's own estimator on the GPU bit for bit
            for key in ("x_scale", "w_scale", "y"):
                assert torch.equal(got[key], stock[key]), f"fused estimator differs from the stock estimator: {key}"
