cd $GRAFT_REPO_ROOT
M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 300 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:'(ew_|bwd_).*bfloat' -c 40 --csv --log-file gpurun_out/r2x_stalls_bf16.csv python tools/bench_extras.py > gpurun_out/r2x_extras.log 2>&1
tail -3 gpurun_out/r2x_extras.log | cut -c1-300
ls -la gpurun_out/r2x*
