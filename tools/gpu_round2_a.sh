set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -80) > gpurun_out/r2a_pytest.log 2>&1
tail -30 gpurun_out/r2a_pytest.log
(timeout 300 python tools/bench_gemm.py --json gpurun_out/r2a_gemm.json 2>&1 | tail -20) > gpurun_out/r2a_gemm.log 2>&1
cat gpurun_out/r2a_gemm.log
(timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -5 gpurun_out/r2a_bench.err)
head -c 3000 gpurun_out/r2a_bench.json
