#!/usr/bin/env python
"""Phases of the fused per-tensor calibration kernel (calq_tensor_kernel) from %globaltimer stamps left by thread 0 of
every CTA (FFQ_CALQ_PROF=1): launch skew, extrema pass, grid barrier, parameters, quantize pass.

    FFQ_CALQ_PROF=1 python tools/prof_calq_phases.py"""
import json
import os
import sys

os.environ["FFQ_CALQ_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fastforward_b200 import ops  # noqa: E402

dev = torch.device("cuda")
for shape in [(2048, 4096), (2048, 14336), (8192, 4096)]:
    x = (torch.randn(shape, device=dev) * 0.02).bfloat16()
    mn = torch.full((1,), float("inf"), dtype=torch.bfloat16, device=dev); mx = -mn
    scale, offset = torch.empty(1, device=dev), torch.empty(1, device=dev)
    ws = torch.zeros(ops._CALQ_WS, dtype=torch.uint8, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        e0.record()
        ops.calibrate_quantize_(mn, mx, x, shape, 8, False, True, scale, offset, flags, None, rowsum=True, workspace=ws)
        e1.record()
    torch.cuda.synchronize()
    st = ws[16 + 4096:].view(torch.int64)[: 448 * 8].view(448, 8).cpu()
    live = st[st[:, 0] > 0].double()
    t0 = live[:, 0].min()
    rel = (live[:, :5] - t0) / 1e3                       # us since the first CTA started
    out = {
        "shape": list(shape), "event_us": round(e0.elapsed_time(e1) * 1e3, 1), "ctas": int(live.shape[0]),
        "start_us_max": round(float(rel[:, 0].max()), 2),
        "extrema_done_us_mean_max": [round(float(rel[:, 1].mean()), 2), round(float(rel[:, 1].max()), 2)],
        "barrier_passed_us_mean_max": [round(float(rel[:, 2].mean()), 2), round(float(rel[:, 2].max()), 2)],
        "parameters_us_mean": round(float(rel[:, 3].mean()), 2),
        "end_us_mean_max": [round(float(rel[:, 4].mean()), 2), round(float(rel[:, 4].max()), 2)],
        "phase_us": {"extrema pass": round(float((rel[:, 1] - rel[:, 0]).mean()), 2),
                     "barrier wait": round(float((rel[:, 2] - rel[:, 1]).mean()), 2),
                     "partials + parameters": round(float((rel[:, 3] - rel[:, 2]).mean()), 2),
                     "quantize pass": round(float((rel[:, 4] - rel[:, 3]).mean()), 2)},
    }
    print(json.dumps(out), flush=True)
