set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2d_pytest.log 2>&1; tail -8 gpurun_out/r2d_pytest.log)
for ds in 0 1; do echo "DIRECT_STORE=$ds"; if [ $ds = 1 ]; then export FFQ_GEMM_DIRECT_STORE=1; else unset FFQ_GEMM_DIRECT_STORE; fi; timeout 200 python tools/bench_gemm.py --clusters 2 2>&1 | cut -c1-220; done > gpurun_out/r2d_gemm.log 2>&1
unset FFQ_GEMM_DIRECT_STORE
cat gpurun_out/r2d_gemm.log
