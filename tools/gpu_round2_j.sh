set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_qlinear_gpu.py tests/test_round2_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2j_pytest.log 2>&1; tail -15 gpurun_out/r2j_pytest.log)
(timeout 300 python tools/bench_gemm.py --wide-tw 512,448,416,384 > gpurun_out/r2j_gemm.log 2>&1; cut -c1-700 gpurun_out/r2j_gemm.log)
(timeout 120 python tools/prof_gemm_roles.py > gpurun_out/r2j_gemm_roles.jsonl 2>&1; cut -c1-1300 gpurun_out/r2j_gemm_roles.jsonl)
