set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x --deselect tests/test_reference_suite_gpu.py > gpurun_out/r2b_pytest.log 2>&1; tail -15 gpurun_out/r2b_pytest.log)
(timeout 600 python -m pytest tests/test_reference_suite_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2b_pytest_ref.log 2>&1; tail -5 gpurun_out/r2b_pytest_ref.log)
timeout 300 python tools/run_ref_tests.py --mode default-cuda --no-plugin --log gpurun_out/ref_tests_default_cuda_noplugin.log -rf > /dev/null 2>&1
tail -3 gpurun_out/ref_tests_default_cuda_noplugin.log
timeout 300 python tools/prof_gemm_roles.py > gpurun_out/r2b_gemm_roles.jsonl 2>&1; cat gpurun_out/r2b_gemm_roles.jsonl
METRICS=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__cycles_active.avg,sm__pipe_tensor_op_imma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpc__cycles_elapsed.avg.per_second
timeout 600 ncu --metrics $METRICS --clock-control none -k regex:'w8a8_gemm2|gemm|cutlass|sm100|xmma' -c 6 --csv --log-file gpurun_out/r2b_ncu_gemm.csv python tools/ncu_gemm.py > gpurun_out/r2b_ncu_gemm.log 2>&1
tail -3 gpurun_out/r2b_ncu_gemm.log; wc -c gpurun_out/r2b_ncu_gemm.csv
(timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -5 gpurun_out/r2b_bench.err)
head -c 1500 gpurun_out/r2b_bench.json
