"""Generate tests/golden/*.pt by running the UNMODIFIED reference on seeded inputs.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference); the
fixtures it writes are committed so that the GPU box, which has no /root/reference, can
check both the oracle (oracle/ref_ops.py) and the CUDA path against real reference output.

    python oracle/make_golden.py            # rewrites tests/golden/*.pt

The reference imports after shimming two absent, off-path dependencies:
``optree`` (only ``tree_map`` is used, quantized_tensor.py:561-562) and the libcst/mypy
based ``fastforward.autoquant`` modules.
"""

from __future__ import annotations

import gzip
import itertools
import os
import sys
import types

import torch

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _import_reference():
    import torch.utils._pytree as _pt

    optree = types.ModuleType("optree")
    optree.tree_map = lambda fn, tree, *rest, **kw: _pt.tree_map(fn, tree)
    sys.modules.setdefault("optree", optree)

    def _na(*a, **k):
        raise NotImplementedError("autoquant is stubbed")

    for name, attrs in (("fastforward.autoquant", {"autoquantize": _na}), ("fastforward.testing.autoquant", {})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    sys.path.insert(0, REF_SRC)
    import fastforward as ff  # noqa

    return ff


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def static_cases(ff):
    """torch.ops.fastforward.{quantize,dequantize,quantize_by_tile_backward} on a layout x dtype matrix.

    Layouts follow the coverage the reference's cuda-parametrised tests demand
    (tests/quantization/test_tiled_affine.py:34-40,58-261)."""
    ops = torch.ops.fastforward
    layouts = [
        # (shape, tile, full dtype matrix?)
        ((8, 4, 2), (8, 4, 2), True),        # per tensor
        ((16, 8, 4), (1, 8, 4), True),       # per channel 0
        ((16, 8, 4), (16, 1, 4), False),     # per channel 1
        ((16, 8, 4), (16, 8, 1), True),      # per channel 2 (strided)
        ((16, 8, 4), (8, 1, 4), False),      # block (channel 1, block axis 0, 8)
        ((16, 8, 4), (4, 8, 1), False),      # block (channel 2, axis 0, 4)
        ((16, 8, 4), (1, 4, 4), False),      # block (channel 0, axis 1, 4)
        ((16, 8, 4), (1, 8, 2), False),      # block (channel 0, axis 2, 2)
        ((16, 8, 4), (8, 4, 2), True),       # arbitrary tile
        ((16, 8, 4), (4, 8, 1), False),
        ((8, 4, 2), (1, 1, 1), False),       # per element
        ((8, 256), (1, 256), True),          # weight per-channel
        ((8, 256), (1, 128), True),          # weight per-group g=128
        ((12, 96), (1, 32), False),          # g=32
        ((16, 64), (16, 1), True),           # per channel last dim
        ((5, 7), (5, 7), False),             # odd sizes, per tensor (no vector alignment)
        ((6, 35), (1, 35), False),           # odd row length
        ((2, 2048), (1, 2048), False),
        ((1, 4096), (1, 4096), False),
    ]
    cases = []
    seed = 100
    for (shape, tile, full), num_bits in itertools.product(layouts, [2, 3, 4, 8]):
        ntiles = 1
        for d, t in zip(shape, tile):
            ntiles *= d // t
        dtype_matrix = [
            # (data dtype, scale dtype, offset dtype or None, output dtype)
            (torch.float32, torch.float32, torch.float32, torch.float32),
            (torch.float32, torch.float32, None, torch.float32),
        ]
        if full and num_bits in (4, 8):
            dtype_matrix += [
                (torch.float32, torch.float32, torch.float32, torch.int32),
                (torch.float32, torch.float32, torch.float32, torch.int8),
                (torch.bfloat16, torch.float32, torch.float32, torch.bfloat16),
                (torch.bfloat16, torch.bfloat16, torch.bfloat16, torch.bfloat16),
                (torch.float16, torch.float16, torch.float16, torch.float16),
                (torch.float16, torch.float32, None, torch.int16),
                (torch.bfloat16, torch.float32, None, torch.int8),
            ]
        if full and num_bits == 3:
            dtype_matrix += [
                (torch.float32, torch.float16, torch.float16, torch.float32),
                (torch.float32, torch.float32, torch.int32, torch.int16),
                (torch.int32, torch.float32, torch.float32, torch.float32),
                (torch.int16, torch.float32, torch.int8, torch.int32),
                (torch.int8, torch.float16, torch.int16, torch.float32),
            ]
        for xdt, sdt, odt, qdt in dtype_matrix:
            seed += 1
            g = _gen(seed)
            if xdt.is_floating_point:
                x = (torch.randn(shape, generator=g) * 1.7).to(xdt)
            else:
                x = torch.randint(-20, 20, shape, generator=g).to(xdt)
            # scales chosen so that a good share of elements clip and ties occur
            scale = (torch.rand(ntiles, generator=g) * 0.4 + 0.05) * (16.0 / 2 ** num_bits + 0.2)
            if not xdt.is_floating_point:
                scale = scale * 6
            scale = scale.to(sdt)
            offset = None
            if odt is not None:
                offset = (torch.randn(ntiles, generator=g) * (2 ** num_bits) * 0.2)
                offset = offset.round().to(odt) if not odt.is_floating_point else offset.to(odt)
            grad = torch.randn(shape, generator=g)
            grad = grad.to(xdt) if xdt.is_floating_point else grad.to(sdt)

            q = ops.quantize_by_tile(x, scale, list(tile), float(num_bits), qdt, offset)
            ddt = xdt if xdt.is_floating_point else sdt
            y = ops.dequantize_by_tile(q, scale, list(tile), offset, ddt)
            case = dict(
                kind="static", shape=shape, tile=tile, num_bits=num_bits,
                x=x, scale=scale, offset=offset, grad=grad, qdtype=qdt, ddtype=ddt, q=q, y=y,
            )
            if xdt.is_floating_point:
                dx, dscale, doffset = ops.quantize_by_tile_backward(
                    x, grad, scale, list(tile), float(num_bits), offset
                )
                case.update(dx=dx, dscale=dscale, doffset=doffset)
            cases.append(case)
    return cases


def quantizer_cases(ff):
    """LinearQuantizer through the public API: range -> params -> codes -> dequant -> autograd."""
    cases = []
    seed = 5000
    grans = [
        ("per_tensor", lambda: ff.PerTensor(), (16, 64)),
        ("per_channel0", lambda: ff.PerChannel(0), (16, 64)),
        ("per_channel1", lambda: ff.PerChannel(1), (16, 64)),
        ("per_block128", lambda: ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0), (8, 256)),
        ("per_block32", lambda: ff.PerBlock(block_dims=1, block_sizes=32, per_channel_dims=0), (8, 96)),
        ("per_tile", lambda: ff.PerTile((4, 16)), (16, 64)),
    ]
    for (gname, gfn, shape), num_bits, symmetric, allow_one_sided, positive, xdt in itertools.product(
        grans, [4, 8], [True, False], [True, False], [False, True], [torch.float32, torch.bfloat16]
    ):
        if xdt is torch.bfloat16 and (num_bits == 4 and not symmetric):
            continue
        seed += 1
        g = _gen(seed)
        x = torch.randn(shape, generator=g)
        if positive:
            x = x.abs() + 0.01
        x = x.to(xdt)
        quantizer = ff.nn.LinearQuantizer(
            num_bits, symmetric=symmetric, allow_one_sided=allow_one_sided, granularity=gfn()
        )
        tile = quantizer.granularity.tile_size(x.shape)
        tile = tuple(x.shape) if tile == "data_shape" else tuple(tile)
        rows = ff.quantization.tiled_tensor.tiles_to_rows(x, tile)
        rmin, rmax = rows.min(-1).values * 0.8, rows.max(-1).values * 0.8   # 0.8: force clipping
        quantizer.quantization_range = (rmin, rmax)
        xg = x.clone().requires_grad_(True)
        qt = quantizer(xg)
        y = qt.dequantize()
        grad = torch.randn(shape, generator=g).to(xdt)
        y.backward(grad)
        offset = quantizer.offset
        cases.append(dict(
            kind="quantizer", gran=gname, shape=shape, tile=tile, num_bits=num_bits, symmetric=symmetric,
            allow_one_sided=allow_one_sided, x=x, range_min=rmin, range_max=rmax, grad=grad,
            scale=quantizer.scale.detach().clone(),
            offset=None if offset is None else offset.detach().clone(),
            offset_is_param=isinstance(offset, torch.nn.Parameter),
            q=qt.raw_data.detach().clone(), y=y.detach().clone(), dx=xg.grad.clone(),
            dscale=quantizer.scale.grad.clone(),
            doffset=None if not isinstance(offset, torch.nn.Parameter) else offset.grad.clone(),
            range_after=tuple(t.detach().clone() for t in quantizer.quantization_range),
        ))
    return cases


def minmax_cases(ff):
    """ff.estimate_ranges(model, running_minmax) over 5 batches; final ranges + params + last codes."""
    cases = []
    seed = 9000
    for gname, gfn, shape in [
        ("per_tensor", lambda: ff.PerTensor(), (2, 16, 24)),
        ("per_channel_last", lambda: ff.PerChannel(2), (2, 16, 24)),
        ("per_channel0", lambda: ff.PerChannel(0), (24, 64)),
        ("per_block", lambda: ff.PerBlock(block_dims=1, block_sizes=16, per_channel_dims=0), (24, 64)),
    ]:
        for symmetric, disable_q, xdt in itertools.product([True, False], [False, True], [torch.float32, torch.bfloat16]):
            seed += 1
            g = _gen(seed)
            quantizer = ff.nn.LinearQuantizer(8, symmetric=symmetric, granularity=gfn())
            batches = [(torch.randn(shape, generator=g) * (i + 1) * 0.3 + 0.1 * i).to(xdt) for i in range(5)]
            outs = []
            with torch.no_grad(), ff.estimate_ranges(quantizer, ff.range_setting.running_minmax,
                                                     disable_quantization=disable_q):
                for b in batches:
                    o = quantizer(b)
                    outs.append(o)
            last = outs[-1]
            last_raw = last.raw_data.clone() if isinstance(last, ff.QuantizedTensor) else last.clone()
            rng = tuple(t.detach().clone() for t in quantizer.quantization_range)
            cases.append(dict(
                kind="running_minmax", gran=gname, shape=shape, symmetric=symmetric, disable_quantization=disable_q,
                batches=batches, scale=quantizer.scale.detach().clone(),
                offset=None if quantizer.offset is None else quantizer.offset.detach().clone(),
                last_raw=last_raw, range=rng,
            ))
    return cases


def calib_int8_cases(ff):
    """The W8A8 calibration recipe: LinearQuantizer(quantized_dtype=int8) under estimate_ranges(running_minmax) over
    3 batches, on layouts the fused calibration kernels cover (per-channel rows, whole tensors).  Records the codes
    of EVERY step (each step quantizes with the range known so far), the final parameters and the running range.
    Data variants exercise the global one-sided decision: mixed-sign, all rows non-negative, some rows non-negative."""
    cases = []
    seed = 12000
    for gname, gfn, shape in [
        ("per_tensor", lambda: ff.PerTensor(), (6, 640)),
        ("per_channel0", lambda: ff.PerChannel(0), (5, 1024)),
    ]:
        for (symmetric, one_sided), xdt, variant, bits in itertools.product(
                [(True, True), (True, False), (False, True)], [torch.float32, torch.bfloat16],
                ["mixed", "positive", "some_positive"], [8, 4]):
            if bits == 4 and (variant != "mixed" or xdt != torch.bfloat16):
                continue
            seed += 1
            g = _gen(seed)
            quantizer = ff.nn.LinearQuantizer(bits, symmetric=symmetric, allow_one_sided=one_sided, granularity=gfn(),
                                              quantized_dtype=torch.int8)
            batches = []
            for i in range(3):
                b = torch.randn(shape, generator=g) * (0.4 + 0.3 * i)
                if variant == "positive":
                    b = b.abs() + 0.01
                elif variant == "some_positive":
                    b[::2] = b[::2].abs()
                batches.append(b.to(xdt))
            raws = []
            with torch.no_grad(), ff.estimate_ranges(quantizer, ff.range_setting.running_minmax):
                for b in batches:
                    raws.append(quantizer(b).raw_data.clone())
            rng = tuple(t.detach().clone() for t in quantizer.quantization_range)
            cases.append(dict(
                kind="calib_int8", gran=gname, shape=shape, symmetric=symmetric, allow_one_sided=one_sided, num_bits=bits,
                variant=variant, batches=batches, raws=raws, scale=quantizer.scale.detach().clone(),
                offset=None if quantizer.offset is None else quantizer.offset.detach().clone(), range=rng))
    return cases


def gptq_cases(ff):
    """fastforward.quantization.gptq.gptq() on seeded QuantizedLinear layers: initial weight, calibration inputs,
    the calibrated parameters BEFORE the block loop (smoothed_minmax on the fp32 weight), final weight and
    final parameters (group scales are recomputed for PerBlock without actorder)."""
    import importlib
    ref_gptq = importlib.import_module('fastforward.quantization.gptq')

    cases = []
    seed = 15000
    grans = {
        "per_tensor": lambda: ff.PerTensor(),
        "per_channel0": lambda: ff.PerChannel(0),
        "per_channel1": lambda: ff.PerChannel(1),
        "per_block32": lambda: ff.PerBlock(block_dims=1, block_sizes=32, per_channel_dims=0),
        "per_tile": lambda: ff.PerTile((4, 48)),
    }
    for gname, bits, symmetric, actorder, block_size, qdtype in [
        ("per_channel0", 4, True, False, 64, None), ("per_channel0", 4, False, True, 64, None),
        ("per_tensor", 8, True, False, 128, torch.int8), ("per_channel1", 4, False, False, 32, None),
        ("per_block32", 4, True, False, 64, None), ("per_block32", 3, False, False, 48, torch.int8),
        ("per_block32", 4, True, True, 64, None), ("per_tile", 4, False, False, 128, None),
        ("per_channel0", 2, True, False, 128, None),
    ]:
        seed += 1
        g = _gen(seed)
        rows, cols = 24, 192
        layer = ff.nn.QuantizedLinear(cols, rows, bias=False)
        with torch.no_grad():
            layer.weight.copy_(torch.randn(rows, cols, generator=g) * 0.05)
        layer.weight_quantizer = ff.nn.LinearQuantizer(bits, symmetric=symmetric, granularity=grans[gname](),
                                                       quantized_dtype=qdtype)
        acts = [torch.randn(2, 40, cols, generator=g) * (1.0 + 0.5 * torch.rand(cols, generator=g)) for _ in range(3)]
        w0 = layer.weight.detach().clone()
        # the parameters the block loop starts from: what gptq() computes first (gptq.py:76-77)
        probe = ff.nn.LinearQuantizer(bits, symmetric=symmetric, granularity=grans[gname](), quantized_dtype=qdtype)
        with ff.estimate_ranges(probe, ff.range_setting.smoothed_minmax):
            probe(w0.clone().float())
        dataset = [((a.clone(),), {}) for a in acts]
        with torch.no_grad():
            ref_gptq.gptq(layer, dataset, block_size=block_size, perc_damp=0.01, actorder=actorder)
        wq = layer.weight_quantizer
        cases.append(dict(
            kind="gptq", gran=gname, num_bits=bits, symmetric=symmetric, actorder=actorder, block_size=block_size,
            qdtype=qdtype, weight=w0, activations=acts, scale0=probe.scale.detach().clone(),
            offset0=None if probe.offset is None else probe.offset.detach().clone(),
            tile=tuple(wq.granularity.tile_size(w0.shape)) if not isinstance(wq.granularity.tile_size(w0.shape), str) else tuple(w0.shape),
            new_weight=layer.weight.detach().clone(), scale=wq.scale.detach().clone(),
            offset=None if wq.offset is None else wq.offset.detach().clone()))
    return cases


def dynamic_cases(ff):
    ops = torch.ops.fastforward
    cases = []
    seed = 12000
    for (shape, tile), num_bits, symmetric, allow_one_sided, positive, xdt in itertools.product(
        [((16, 64), (16, 64)), ((16, 64), (1, 64)), ((16, 64), (1, 16)), ((16, 64), (16, 1)), ((8, 6, 10), (2, 3, 5))],
        [4, 8], [True, False], [True, False], [False, True], [torch.float32, torch.bfloat16],
    ):
        seed += 1
        g = _gen(seed)
        x = torch.randn(shape, generator=g)
        if positive:
            x = x.abs()
        x = x.to(xdt)
        qdt = xdt
        q, scale, offset = ops.quantize_dynamic_by_tile(x, list(tile), float(num_bits), symmetric, allow_one_sided, qdt)
        cases.append(dict(kind="dynamic", shape=shape, tile=tile, num_bits=num_bits, symmetric=symmetric,
                          allow_one_sided=allow_one_sided, x=x, qdtype=qdt, q=q, scale=scale, offset=offset))
    return cases


def linear_cases(ff):
    """QuantizedLinear through ff.nn.functional.linear -> dequantize fallback (_gen/fallback.py:77-112)."""
    cases = []
    seed = 15000
    for (m, k, n), bias, wsym, xdt in itertools.product(
        [(32, 64, 48), (64, 128, 32), (17, 96, 40)], [False, True], [True, False], [torch.float32, torch.bfloat16]
    ):
        seed += 1
        g = _gen(seed)
        x = torch.randn(m, k, generator=g).to(xdt)
        w = (torch.randn(n, k, generator=g) * 0.05).to(xdt)
        b = (torch.randn(n, generator=g) * 0.1).to(xdt) if bias else None
        xq = ff.nn.LinearQuantizer(8, symmetric=False, granularity=ff.PerTensor(), quantized_dtype=torch.int8)
        wq = ff.nn.LinearQuantizer(8, symmetric=wsym, granularity=ff.PerChannel(0), quantized_dtype=torch.int8)
        xq.quantization_range = (x.min(), x.max())
        wq.quantization_range = (w.min(1).values, w.max(1).values)
        with torch.no_grad():
            xqt, wqt = xq(x), wq(w)
            y = ff.nn.functional.linear(xqt, wqt, b, output_quantizer=None, strict_quantization=False)
        cases.append(dict(
            kind="linear", m=m, k=k, n=n, x=x, w=w, bias=b,
            x_codes=xqt.raw_data.clone(), x_scale=xq.scale.detach().clone(), x_offset=xq.offset.detach().clone(),
            w_codes=wqt.raw_data.clone(), w_scale=wq.scale.detach().clone(),
            w_offset=None if wq.offset is None else wq.offset.detach().clone(), y=y,
        ))
    return cases


def mse_grid_cases(ff):
    """ff.estimate_ranges(quantizer, mse_grid, num_candidates=C) over 3 batches: the search grid,
    the accumulated errors and the parameters the estimator leaves in the quantizer
    (range_setting/min_error.py:64-221; tests/range_setting/test_minerror.py)."""
    cases = []
    seed = 18000
    for (gname, gfn, shape), symmetric, positive, xdt, ncand in itertools.product(
        [("per_tensor", lambda: ff.PerTensor(), (4, 16, 32)),
         ("per_channel_last", lambda: ff.PerChannel(2), (2, 16, 24)),
         ("per_channel0", lambda: ff.PerChannel(0), (24, 64)),
         ("per_block", lambda: ff.PerBlock(block_dims=1, block_sizes=16, per_channel_dims=0), (24, 64))],
        [True, False], [False, True], [torch.float32, torch.bfloat16], [7, 20],
    ):
        seed += 1
        g = _gen(seed)
        quantizer = ff.nn.LinearQuantizer(4, symmetric=symmetric, granularity=gfn())
        batches = []
        for i in range(3):
            b = torch.randn(shape, generator=g) * (1.0 + 0.25 * i)
            batches.append((b.abs() if positive else b).to(xdt))
        from fastforward.range_setting.min_error import mse_grid

        estimator = mse_grid(num_candidates=ncand)
        with torch.no_grad(), ff.estimate_ranges(quantizer, estimator):
            for b in batches:
                quantizer(b)
            step = next(iter(quantizer._quantizer_overrides.values()))
            grid = (step.min_threshold.clone(), step.max_threshold.clone())
            cumulative = step.cumulative_error.clone()
        cases.append(dict(
            kind="mse_grid", gran=gname, shape=shape, symmetric=symmetric, positive=positive, num_candidates=ncand,
            num_bits=4, batches=batches, min_threshold=grid[0], max_threshold=grid[1], cumulative_error=cumulative,
            scale=quantizer.scale.detach().clone(),
            offset=None if quantizer.offset is None else quantizer.offset.detach().clone(),
        ))
    return cases


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)   # reduction order of aten sum is thread-count dependent only above the grain size
    ff = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    for name, fn in [
        ("static", static_cases), ("quantizer", quantizer_cases), ("running_minmax", minmax_cases),
        ("dynamic", dynamic_cases), ("linear", linear_cases), ("mse_grid", mse_grid_cases),
        ("calib_int8", calib_int8_cases), ("gptq", gptq_cases),
    ]:
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue            # `python oracle/make_golden.py mse_grid` regenerates one fixture only
        cases = fn(ff)
        path = os.path.join(OUT, f"{name}.pt.gz")
        with gzip.open(path, "wb", compresslevel=9) as fh:
            torch.save(dict(torch=torch.__version__, generator="oracle/make_golden.py", cases=cases), fh)
        print(f"{name}: {len(cases)} cases -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
