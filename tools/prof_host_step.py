#!/usr/bin/env python
"""Host-side profile (cProfile) of an EAGER calibration step of the bench workload: where the Python time of
`with ff.estimate_ranges(...)` goes when no CUDA graph hides it.

    python tools/prof_host_step.py [layers]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench_workloads as bw  # noqa: E402
import fastforward_b200 as ff  # noqa: E402
from fastforward_b200.nn import qlinear  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda")
model = bw.DecoderStack(bw.LLAMA3_8B, layers=layers, dtype=torch.bfloat16, device=dev)
bw.init_weights_(model)
bw.quantize_for_w8a8(ff, model)
model.to(dev)
qlinear.install()
tokens = torch.randint(0, 1000, (1, 2048), device=dev)
with torch.no_grad(), ff.estimate_ranges(model, ff.range_setting.running_minmax(memoize_parameters=False)):
    for _ in range(3):
        model(tokens)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        model(tokens)
    t_host = (time.perf_counter() - t0) / 5
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 5
    print(f"{layers} layers: host issue time {t_host * 1e3:.2f} ms per step, with the device {t_all * 1e3:.2f} ms "
          f"({t_host * 1e6 / (7 * layers):.0f} us of host time per quantized linear)")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        model(tokens)
    pr.disable()
    torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(30)
