#!/usr/bin/env python
"""A few eager calibration steps of ONE decoder layer of the bench workload (Llama-3-8B shapes, W8A8 recipe), for `ncu`:

    ncu --set full --clock-control none --import-source on -k regex:'ffq|calq|w8a8' -o gpurun_out/layer \\
        python tools/prof_layer.py [steps]
    ncu -i gpurun_out/layer.ncu-rep --page raw --csv --print-units base > gpurun_out/layer_raw.csv
    python tools/ncu_layer_summary.py gpurun_out/layer_raw.csv profiles/r02_ncu_full_layer.md profiles/r02_traffic.json

Nothing here is a timing: the kernels of the LAST step are the steady state (ranges settled, parameters materialised)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench_workloads as bw  # noqa: E402
import fastforward_b200 as ff  # noqa: E402
from fastforward_b200.nn import qlinear  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda")
model = bw.DecoderStack(bw.LLAMA3_8B, layers=1, dtype=torch.bfloat16, device=dev)
bw.init_weights_(model)
bw.quantize_for_w8a8(ff, model)
model.to(dev)
qlinear.install()
g = torch.Generator().manual_seed(0)
with torch.no_grad(), ff.estimate_ranges(model, ff.range_setting.running_minmax(memoize_parameters=False)):
    for _ in range(steps):
        model(torch.randint(0, bw.LLAMA3_8B.vocab, (1, 2048), generator=g).to(dev))
torch.cuda.synchronize()
print("launches of this library:", ff._cabi.launch_count())
