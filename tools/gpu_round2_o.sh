cd $GRAFT_REPO_ROOT
(timeout 300 python -m pytest tests/test_qlinear_gpu.py tests/test_round2_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3)
timeout 200 python tools/bench_gemm.py 2>&1 | cut -c1-250
timeout 100 python tools/prof_gemm_roles.py 2>&1 | cut -c1-1000
timeout 100 python tools/bench_extras.py 2>&1 | tail -5 | cut -c1-600
