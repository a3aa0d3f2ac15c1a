#!/bin/bash
# round 2, session 3: optimistic per-tensor calibration pair -- parity tests, micro-benchmark A/B, short bench A/B, LPBQ tests
cd "$(dirname "$0")/.."
python -m pytest tests/test_calibrate_gpu.py tests/test_round2_gpu.py tests/test_api_gpu.py -x -q 2>&1 | tail -12 > gpurun_out/r3b_pytest.log
tail -3 gpurun_out/r3b_pytest.log
(cd tools && python bench_calq.py) > gpurun_out/r3b_calq_opt.log 2>&1
(cd tools && FFQ_CALQ_OPTIMISTIC=0 python bench_calq.py) > gpurun_out/r3b_calq_twopass.log 2>&1
grep -h "2048\|8192" gpurun_out/r3b_calq_opt.log | sed 's/^/opt  /'
grep -h "2048\|8192" gpurun_out/r3b_calq_twopass.log | sed 's/^/2pass /'
for v in 1 0; do
  FFQ_CALQ_OPTIMISTIC=$v python bench.py --skip-extras --skip-cpu-baseline --skip-drop-in --skip-compiled-baseline --steps 10 --warmup 3 \
    > gpurun_out/r3b_bench_opt$v.json 2> gpurun_out/r3b_bench_opt$v.err
  python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r3b_bench_opt$v.json").read().splitlines() if l.startswith("{")][-1])
print("optimistic=$v", d["ms_per_step"], d["value"], d["gpu_launches_per_step"], {k: v["ms_per_step"] for k, v in d["kernels"].items()})
P
done
