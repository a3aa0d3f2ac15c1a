cd $GRAFT_REPO_ROOT
run() { echo "== $*"; env "$@" timeout 200 python tools/bench_gemm.py --kernels pair,wide --shapes 8192x14336x4096,2048x4096x4096,2048x14336x4096,2048x4096x14336 2>&1 | cut -c1-300; }
run FFQ_GEMM_DEBUG=0 FFQ_GEMM_WIDE_HALVES=1
run FFQ_GEMM_DEBUG=1 FFQ_GEMM_WIDE_HALVES=1
run FFQ_GEMM_DEBUG=2 FFQ_GEMM_WIDE_HALVES=1
run FFQ_GEMM_DEBUG=0 FFQ_GEMM_WIDE_HALVES=1 FFQ_GEMM_WIDE_TW=224
(FFQ_GEMM_KERNEL=wide FFQ_GEMM_WIDE_HALVES=1 timeout 200 python -m pytest tests/test_qlinear_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3)
(FFQ_GEMM_WIDE_HALVES=1 timeout 100 python tools/prof_gemm_roles.py 8192 14336 4096 2>&1 | cut -c1-900)
