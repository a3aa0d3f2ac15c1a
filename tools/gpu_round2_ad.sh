#!/bin/bash
# round 2, session 3: programmatic dependent launch (FFQ_PDL=1) -- parity tests under it, then same-box A/B of the step
cd "$(dirname "$0")/.."
FFQ_PDL=1 python -m pytest tests/test_calibrate_gpu.py tests/test_qlinear_gpu.py tests/test_round2_gpu.py tests/test_api_gpu.py -x -q 2>&1 | tail -6 > gpurun_out/r3d_pytest_pdl.log
tail -3 gpurun_out/r3d_pytest_pdl.log
for v in 0 1 0 1; do
  FFQ_PDL=$v python bench.py --skip-extras --skip-cpu-baseline --skip-drop-in --skip-compiled-baseline --steps 10 --warmup 3 \
    > gpurun_out/r3d_bench_pdl$v.json 2> gpurun_out/r3d_bench_pdl$v.err
  python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r3d_bench_pdl$v.json").read().splitlines() if l.startswith("{")][-1])
print("pdl=$v", d["ms_per_step"], d["value"], {k: x["ms_per_step"] for k, x in d["kernels"].items()}, d["ablation"]["eager_no_cuda_graph"]["ms_per_step"])
P
done
