cd $GRAFT_REPO_ROOT
for lib in build/ab/libffq_prev.so fastforward_b200/lib/libffq_b200.so; do
echo "== $lib"
(cd tools; FFQ_LIB_PATH=$GRAFT_REPO_ROOT/$lib timeout 100 python bench_group.py 2>&1 | tail -6 | cut -c1-150)
done
M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
(cd tools; timeout 300 ncu --metrics $M --clock-control none -k regex:'calq' -c 60 --csv --log-file ../gpurun_out/r2v_stalls_group.csv python bench_group.py > /dev/null 2>&1)
timeout 300 ncu --metrics $M --clock-control none -k regex:'ew_row|bwd_row|ew_tile' --launch-skip 120 -c 150 --csv --log-file gpurun_out/r2v_stalls_extras.csv python tools/bench_extras.py > /dev/null 2>&1
ls -la gpurun_out/r2v_*
