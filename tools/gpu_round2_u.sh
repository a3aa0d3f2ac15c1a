cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(cd tools; timeout 100 python bench_group.py 2>&1 | tail -6)
(timeout 600 python -m pytest tests/test_calibrate_gpu.py tests/test_round2_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3)
M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 400 ncu --metrics $M --clock-control none -k regex:'ffq|calq|ew_|bwd_|mm_row|grid_mse' -c 120 --csv --log-file gpurun_out/r2u_stalls_extras.csv python tools/bench_extras.py > gpurun_out/r2u_extras.log 2>&1
timeout 200 ncu --metrics $M --clock-control none -k regex:'calq' -c 40 --csv --log-file gpurun_out/r2u_stalls_calq.csv python tools/prof_calq.py > /dev/null 2>&1
ls -la gpurun_out/r2u_*
