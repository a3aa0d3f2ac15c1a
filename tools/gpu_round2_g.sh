set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2g_pytest.log 2>&1; tail -6 gpurun_out/r2g_pytest.log)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'w8a8_gemm2' -c 1 -o gpurun_out/r2g_gemm_full python tools/ncu_gemm.py > gpurun_out/r2g_ncu.log 2>&1
tail -3 gpurun_out/r2g_ncu.log
ls -la gpurun_out/*.ncu-rep
