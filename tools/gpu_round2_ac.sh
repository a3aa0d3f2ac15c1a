#!/bin/bash
# round 2, session 3: pair GEMM with eight epilogue warps (default build) against the four-warp build (lib/libffq_b200_ep4.so),
# optimistic per-tensor pair with 8 against 4 loads in flight
cd "$(dirname "$0")/.."
EP4=$PWD/fastforward_b200/lib/libffq_b200_ep4.so
python -m pytest tests/test_qlinear_gpu.py tests/test_round2_gpu.py tests/test_calibrate_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/r3c_pytest.log
tail -3 gpurun_out/r3c_pytest.log
SH=8192x14336x4096,2048x4096x4096,2048x14336x4096,2048x4096x14336,2048x1024x4096
python tools/bench_gemm.py --shapes $SH > gpurun_out/r3c_gemm_ep8.log 2>&1
FFQ_LIB_PATH=$EP4 python tools/bench_gemm.py --shapes $SH > gpurun_out/r3c_gemm_ep4.log 2>&1
python tools/bench_gemm.py --shapes $SH > gpurun_out/r3c_gemm_ep8_b.log 2>&1
for f in ep8 ep4 ep8_b; do echo "== $f"; cut -c1-200 gpurun_out/r3c_gemm_$f.log | tail -6; done
(cd tools && python bench_calq.py) 2>&1 | grep "2048\|8192" | sed 's/^/U8 /'
(cd tools && FFQ_CALQ_OPT_U=4 python bench_calq.py) 2>&1 | grep "2048\|8192" | sed 's/^/U4 /'
for v in ep8 ep4; do
  L=""; [ $v = ep4 ] && L=$EP4
  FFQ_LIB_PATH=$L python bench.py --skip-extras --skip-cpu-baseline --skip-drop-in --skip-compiled-baseline --steps 10 --warmup 3 \
    > gpurun_out/r3c_bench_$v.json 2> gpurun_out/r3c_bench_$v.err
  python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r3c_bench_$v.json").read().splitlines() if l.startswith("{")][-1])
print("$v", d["ms_per_step"], d["value"], {k: x["ms_per_step"] for k, x in d["kernels"].items()}, d["roofline"]["frac"])
P
done
