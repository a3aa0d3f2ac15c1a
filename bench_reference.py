"""Reference arms of bench.py  --  BASELINE INFRASTRUCTURE, never part of the product path.

Everything here drives the UNMODIFIED reference package staged under ``oracle/_ref`` (``oracle/fetch_ref.py``)
through its own public API, exactly as bench.py drives ``fastforward_b200``:

  * ``cpu_calibration``      the reference's CPU eager path on the host cores (``--impl reference`` and the
                             ``cpu_baseline`` object of the default line), on a bounded sample of the workload;
  * ``gpu_calibration``      the reference on the B200: (a) alone -- its aten eager chain on CUDA tensors, the real
                             "beat this" bar --, (b) with ``fastforward_b200.plugin.install()`` underneath (the drop-in:
                             reference host code, our kernels, eager, no CUDA graph), (c) as (b) plus the sync-free
                             ``running_minmax`` estimator;
  * ``cfg1_fake_quant``      BASELINE.json configs[0]: 4096x4096 fp32, 8-bit PerChannel LinearQuantizer,
                             ``q(x).dequantize().backward(g)`` on the CPU, on CUDA eager, with the reference's
                             ``compiled_quant_funcs`` flag (flags.py:96-98), and with the plugin.

Falls back to the oracle port (``oracle/workload.py``) only when ``oracle/_ref`` is absent, and says so in ``kind``."""
from __future__ import annotations

import os
import statistics
import time

import torch


def reference_available() -> bool:
    from oracle import ref_loader

    return ref_loader.available()


def _ref():
    from oracle import ref_loader

    return ref_loader.load_reference()


def _build_model(ff, sh, layers, device, disable_strict=True):
    import bench_workloads as bw

    model = bw.DecoderStack(sh, layers=layers, dtype=torch.bfloat16, device=device)
    bw.init_weights_(model, seed=0)
    bw.quantize_for_w8a8(ff, model)
    model.to(device)                 # the quantizers are created on the CPU (LinearQuantizer's default device)
    return model


# ---------------------------------------------------------------------------------------------------------------
# CPU: the reference's own eager path on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_calibration(sh, seq, full_layers, sample_layers, steps, warmup):
    """Really runs `warmup` + `steps` calibration forwards of a `sample_layers`-layer stack (same per-layer shapes,
    quantizers and batches as the GPU arm) and scales tokens/s by sample_layers/full_layers."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    toks = [torch.randint(0, sh.vocab, (1, seq), generator=g) for _ in range(warmup + steps)]
    if reference_available():
        ff = _ref()
        kind, what = "reference", "the unmodified reference (oracle/_ref) through its public API: quantize_model, " \
                                  "find_quantizers().initialize(LinearQuantizer), estimate_ranges(running_minmax)"
        model = _build_model(ff, sh, sample_layers, "cpu")
        ctx = ff.estimate_ranges(model, ff.range_setting.running_minmax)
    else:
        import contextlib

        import bench_workloads as bw
        from oracle import workload as ow

        kind, what = "port", "oracle/workload.py (torch CPU eager ops in the reference's order; oracle/_ref not staged)"
        model = bw.DecoderStack(sh, layers=sample_layers, dtype=torch.bfloat16, device="cpu")
        bw.init_weights_(model, seed=0)
        ow.oracle_calibration_model(model)
        ctx = contextlib.nullcontext()
    per_step = []
    with torch.no_grad(), ctx:
        for i in range(warmup):
            model(toks[i])
        for i in range(steps):
            t0 = time.perf_counter()
            model(toks[warmup + i])
            per_step.append(time.perf_counter() - t0)
    total = sum(per_step)
    step_s = total / max(steps, 1)
    tok_s = seq / (step_s * full_layers / sample_layers)
    return dict(value=round(tok_s, 2), unit="tokens/s", cores=cores, kind=kind,
                sample=f"{steps} timed step(s) (+{warmup} warm-up), each one calibration forward of {sample_layers} of the "
                       f"{full_layers} decoder layers at seq {seq} (bf16, same quantizers and seeded batches); tokens/s = "
                       f"seq / (step time x {full_layers}/{sample_layers}); {what}",
                seconds=round(total, 2), ms_per_sample_step=round(1e3 * step_s, 1),
                ms_per_sample_step_median=round(1e3 * statistics.median(per_step), 1) if per_step else None,
                sample_fraction=sample_layers / full_layers)


# ---------------------------------------------------------------------------------------------------------------
# GPU: the reference on the B200, alone and with the plugin
# ---------------------------------------------------------------------------------------------------------------
def _time_steps(model, ctx_factory, tokens, steps, warmup):
    """tokens/s of `steps` eager calibration forwards inside one estimate_ranges block (block exit included)."""
    seq = tokens[0].shape[-1]
    with torch.no_grad(), ctx_factory():
        for i in range(warmup):
            model(tokens[i % len(tokens)])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            model(tokens[(warmup + i) % len(tokens)])
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0
    dev_s = e0.elapsed_time(e1) * 1e-3
    return dict(tokens_per_s=round(steps * seq / dev_s, 1), ms_per_step=round(1e3 * dev_s / steps, 2),
                wall_ms_per_step=round(1e3 * wall / steps, 2), steps=steps, warmup=warmup)


def gpu_calibration(sh, seq, layers, dev, steps=3, warmup=2, with_plugin=True, cfg1=None, compiled=True):
    """(a) reference alone on CUDA, (b) reference + plugin kernels, (c) + sync-free estimator.  All eager.
    `cfg1` (a dict) additionally receives BASELINE configs[0] measured the same three ways: plugin.install() cannot be
    undone inside a process, so everything "reference alone" runs first."""
    from fastforward_b200 import _cabi

    ff = _ref()
    out = {"layers": layers, "seq": seq, "cuda_graph": False,
           "what": "the unmodified reference's host code (quantizers, estimator, dispatcher, QuantizedTensor) driving the "
                   "same Llama-shape stack on cuda:0"}
    g = torch.Generator().manual_seed(1234)
    tokens = [torch.randint(0, sh.vocab, (1, seq), generator=g).to(dev) for _ in range(warmup + steps)]
    model = _build_model(ff, sh, layers, dev)

    def reset():
        for m in model.modules():
            if isinstance(m, ff.nn.Quantizer) and hasattr(m, "reset_parameters"):
                m.reset_parameters()

    stock = lambda: ff.estimate_ranges(model, ff.range_setting.running_minmax)      # noqa: E731
    if "quantize_by_tile" not in _plugin_state():
        out["reference_alone_eager_cuda"] = _time_steps(model, stock, tokens, steps, warmup)
        if cfg1 is not None:
            cfg1.update(cfg1_fake_quant(dev, cpu=True, cuda=True, compiled=compiled, with_plugin=False))
    if with_plugin:
        from fastforward_b200 import plugin

        reset()
        plugin.install(ff, patch_estimators=False)
        if cfg1 is not None:
            cfg1.update(cfg1_fake_quant(dev, cpu=False, cuda=False, compiled=False, with_plugin=True))
        l0 = _cabi.launch_count()
        out["reference_plus_plugin"] = _time_steps(model, stock, tokens, steps, warmup)
        out["reference_plus_plugin"]["ffq_launches_per_step"] = round((_cabi.launch_count() - l0) / (steps + warmup), 1)
        out["reference_plus_plugin"]["what"] = ("plugin.install(): CUDA-key kernels under torch.ops.fastforward.* + the W8A8 "
                                                "tcgen05 linear in the reference's dispatcher; the reference's own estimator "
                                                "(2 host syncs per quantizer per forward)")
        reset()
        plugin.install_estimators(ff)
        l0 = _cabi.launch_count()
        ours = lambda: ff.estimate_ranges(model, ff.range_setting.running_minmax)    # noqa: E731
        out["reference_plus_plugin_estimators"] = _time_steps(model, ours, tokens, steps, warmup)
        out["reference_plus_plugin_estimators"]["ffq_launches_per_step"] = round((_cabi.launch_count() - l0) / (steps + warmup), 1)
        out["reference_plus_plugin_estimators"]["what"] = ("as above + install(patch_estimators=True): "
                                                           "ff.range_setting.running_minmax is the sync-free fused estimator")
        plugin.uninstall_estimators(ff)
    del model
    torch.cuda.empty_cache()
    return out


def _plugin_state():
    from fastforward_b200 import plugin

    return plugin._installed


def cfg1_fake_quant(dev, cpu=True, cuda=True, compiled=True, with_plugin=True):
    """BASELINE.json configs[0] through the reference's public API: q(x).dequantize().backward(g)."""
    ff = _ref()
    out = {"shape": [4096, 4096], "dtype": "fp32", "quantizer": "LinearQuantizer(8, granularity=PerChannel(0))",
           "step": "q(x).dequantize().backward(g), range from exact row min/max"}

    def make(device):
        torch.manual_seed(0)
        x = torch.randn(4096, 4096, device=device, requires_grad=True)
        g = torch.randn(4096, 4096, device=device)
        q = ff.nn.LinearQuantizer(8, granularity=ff.PerChannel(0)).to(device)
        with torch.no_grad():
            q.quantization_range = (x.min(1).values, x.max(1).values)

        def step():
            x.grad = None
            q.scale.grad = None
            q(x).dequantize().backward(g)
        return step

    def time_cpu(step, iters=5):
        step()
        ts = []
        for _ in range(iters):
            t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
        return statistics.median(ts)

    def time_cuda(step, iters=20, warm=3):
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / iters

    by = 5 * 4096 * 4096 * 4
    if cpu:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        t = time_cpu(make("cpu"))
        out["reference_cpu"] = {"ms": round(t * 1e3, 1), "GBps_algorithmic": round(by / t / 1e9, 2), "cores": cores}
    if cuda and "quantize_by_tile" not in _plugin_state():
        t = time_cuda(make(dev))
        out["reference_eager_cuda"] = {"us": round(t * 1e6, 1), "GBps_algorithmic": round(by / t / 1e9, 1)}
        if compiled:
            try:
                with ff.flags.compiled_quant_funcs(True):
                    step = make(dev)
                    t = time_cuda(step, iters=20, warm=4)
                out["reference_compiled_quant_funcs_cuda"] = {"us": round(t * 1e6, 1), "GBps_algorithmic": round(by / t / 1e9, 1),
                                                              "what": "ff.compiled_quant_funcs(True): torch.compile of the op bodies (flags.py:96-98)"}
            except Exception as e:  # noqa: BLE001  (a baseline that cannot be built is reported, not fatal)
                out["reference_compiled_quant_funcs_cuda"] = f"unavailable: {type(e).__name__}: {str(e)[:160]}"
    if with_plugin:
        from fastforward_b200 import plugin

        plugin.install(ff)
        t = time_cuda(make(dev))
        out["reference_plus_plugin_cuda"] = {"us": round(t * 1e6, 1), "GBps_algorithmic": round(by / t / 1e9, 1),
                                             "what": "same reference code, kernels from plugin.install(); includes the reference's "
                                                     "Python dispatch (custom_op + autograd.Function + QuantizedTensor) per call"}
    return out
