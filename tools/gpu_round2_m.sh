cd $GRAFT_REPO_ROOT
run() { echo "== $*"; env "$@" timeout 200 python tools/bench_gemm.py --kernels pair,wide --shapes 8192x14336x4096,2048x4096x4096,2048x14336x4096 2>&1 | cut -c1-300; }
run FFQ_GEMM_DEBUG=1 FFQ_GEMM_WIDE_GROUP=1
run FFQ_GEMM_DEBUG=1 FFQ_GEMM_WIDE_GROUP=2
run FFQ_GEMM_DEBUG=1 FFQ_GEMM_WIDE_GROUP=4
run FFQ_GEMM_DEBUG=0 FFQ_GEMM_WIDE_GROUP=4
(timeout 200 python -m pytest tests/test_qlinear_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3)
