#!/usr/bin/env python
"""Golden fixtures for the on-disk formats  --  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_formats.py          # needs /root/reference (this container), CPU only

Runs the UNMODIFIED reference (imported from /root/reference/src through the same two-module shim as
oracle/make_golden.py) and records

* ``tests/golden/formats/ref_state/``     what ``save_quantization_state`` of the reference writes for a small
  calibrated model (config.yaml + model.safetensors) and ``ref_artifact/`` what ``save_quantized_model`` writes
  (config.yaml, quantizer_state.safetensors, weights.safetensors, manifest.json); ``expect.pt`` holds the tensors and
  quantizer settings a loader must end up with;
* ``tests/golden/lpbq.pt.gz``             inputs and outputs of ``LPBQProcessor.grouped_dynamic_quantize`` /
  ``generate_lpbq_encoding`` (export/_lpbq.py:78-160) on seeded scale tensors in both block orientations.

The files are a few KB; tests/test_save_load.py loads them with this repository's package, and -- wherever
``oracle/_ref`` is staged -- also lets the reference read what this package wrote."""
from __future__ import annotations

import gzip
import importlib
import os
import shutil
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.make_golden import REF_SRC, _import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
FORMATS = os.path.join(OUT, "formats")


def build_model(ff, seed: int = 11) -> torch.nn.Module:
    """Two linears sharing nothing + one INPUT quantizer instance attached twice (a shared quantizer)."""
    torch.manual_seed(seed)
    model = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.Linear(8, 4))
    ff.quantize_model(model)
    ff.find_quantizers(model, "**/[quantizer:parameter/weight]").initialize(
        ff.nn.LinearQuantizer, num_bits=4, granularity=ff.PerBlock(block_dims=1, block_sizes=4, per_channel_dims=0))
    ff.find_quantizers(model, "**/[quantizer:activation/input]").initialize(
        ff.nn.LinearQuantizer, num_bits=8, symmetric=False, granularity=ff.PerTensor(), quantized_dtype=torch.int8)
    ff.find_quantizers(model, "**/[quantizer:activation/output]").initialize(
        ff.nn.LinearQuantizer, num_bits=8, symmetric=True, allow_one_sided=False, granularity=ff.PerChannel(-1))
    return model


def state_fixtures(ff) -> None:
    from fastforward.quantization import save_load as SL

    model = build_model(ff)
    with torch.no_grad(), ff.strict_quantization(False), ff.estimate_ranges(model, ff.range_setting.running_minmax):
        model(torch.randn(5, 16))
    if os.path.isdir(FORMATS):
        shutil.rmtree(FORMATS)
    os.makedirs(FORMATS)
    cfg = SL.save_quantization_state(model, name_or_path="toy/model", cache_dir=os.path.join(FORMATS, "cache"))
    shutil.move(str(cfg.parent), os.path.join(FORMATS, "ref_state"))
    shutil.rmtree(os.path.join(FORMATS, "cache"))
    SL.save_quantized_model(model, os.path.join(FORMATS, "ref_artifact"), name_or_path="toy/model")
    expect = {
        "weights": {k: v.detach().clone() for k, v in model.state_dict().items()},
        "quantizers": {
            name: {"num_bits": q.num_bits, "symmetric": q.symmetric, "allow_one_sided": q.allow_one_sided,
                   "granularity": repr(q.granularity), "quantized_dtype": q.quantized_dtype,
                   "scale": q.scale.detach().clone(), "offset": None if q.offset is None else q.offset.detach().clone()}
            for name, q in ff.nn.quantized_module.named_quantizers(model)},
    }
    torch.save(expect, os.path.join(FORMATS, "expect.pt"))
    n = sum(len(fs) for _, _, fs in os.walk(FORMATS))
    print(f"formats: {n} files under {FORMATS}")


def _load_lpbq():
    """export/_lpbq.py without executing export/__init__.py (which pulls the ONNX pipeline)."""
    pkg = types.ModuleType("fastforward.export")
    pkg.__path__ = [os.path.join(REF_SRC, "fastforward", "export")]
    sys.modules["fastforward.export"] = pkg
    return importlib.import_module("fastforward.export._lpbq"), importlib.import_module("fastforward.export._export_types")


def lpbq_cases(ff) -> list:
    lpbq, types_ = _load_lpbq()
    cases = []
    g = torch.Generator().manual_seed(23)
    for (rows, cols), block, orient in [((16, 32), 4, "rows"), ((64, 128), 16, "rows"), ((24, 40), 8, "cols"),
                                        ((128, 256), 32, "cols"), ((7, 128), 64, "rows")]:
        for cbw, dbw in ((4, 8), (3, 8), (6, 16)):
            if orient == "rows":       # PerBlock(block_dims=1, per_channel_dims=0): scale_2d = [out, in / block]
                shape2d, tile, grouping = (rows, cols // block), (1, block), [1, -1]
            else:                      # PerBlock(block_dims=0, per_channel_dims=1): scale_2d = [rows / block, in]
                shape2d, tile, grouping = (rows // block, cols), (block, 1), [-1, 1]
            scale = torch.rand(shape2d, generator=g) * torch.logspace(-3, 0, shape2d[1]).reshape(1, -1) + 1e-6
            proc = lpbq.LPBQProcessor(compressed_bw=cbw, decompressed_bw=dbw)
            q, f = proc.grouped_dynamic_quantize(scale, grouping, cbw)
            params = types_.ProcessedQuantParams(
                scale=scale.reshape(-1), offset=torch.zeros(scale.numel()), qnn_offset=torch.zeros(scale.numel()),
                bitwidth=cbw, is_symmetric=True, data_shape=torch.Size((rows, cols)), tile_size=torch.Size(tile))
            enc = proc.generate_lpbq_encoding("w", params)
            cases.append(dict(data_shape=(rows, cols), tile_size=tile, orientation=orient, compressed_bw=cbw,
                              decompressed_bw=dbw, scale=scale.reshape(-1).clone(), grouping=grouping,
                              int_scale=q.to(torch.int64).clone(), float_scale=f.clone(), encoding=enc))
    return cases


def main() -> None:
    torch.manual_seed(0)
    torch.set_num_threads(1)
    ff = _import_reference()
    state_fixtures(ff)
    cases = lpbq_cases(ff)
    path = os.path.join(OUT, "lpbq.pt.gz")
    with gzip.open(path, "wb", compresslevel=9) as fh:
        torch.save(dict(torch=torch.__version__, generator="oracle/make_golden_formats.py", cases=cases), fh)
    print(f"lpbq: {len(cases)} cases -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
