cd $GRAFT_REPO_ROOT
for dbg in 1 2; do
echo "FFQ_GEMM_DEBUG=$dbg"
FFQ_GEMM_DEBUG=$dbg timeout 200 python tools/bench_gemm.py --kernels pair,wide 2>&1 | cut -c1-400
done
FFQ_GEMM_DEBUG=1 timeout 100 python tools/prof_gemm_roles.py 8192 14336 4096 2>&1 | cut -c1-900
