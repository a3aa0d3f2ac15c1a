set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1; tail -25 gpurun_out/r2c_pytest.log)
for dbg in 0 1 2 4 6; do echo "FFQ_GEMM_DEBUG=$dbg"; FFQ_GEMM_DEBUG=$dbg timeout 120 python tools/bench_gemm.py --clusters 2 2>&1 | grep -E "8192x14336x4096|2048x4096x4096 |2048x14336x4096" | cut -c1-200; done > gpurun_out/r2c_gemm_debug.log 2>&1
cat gpurun_out/r2c_gemm_debug.log
(timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -5 gpurun_out/r2c_bench.err)
head -c 1200 gpurun_out/r2c_bench.json
