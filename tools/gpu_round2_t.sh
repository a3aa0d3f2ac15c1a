cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:calq_tile_thread -c 1 --launch-skip 2 -o gpurun_out/r2t_group_full -f python tools/prof_group.py > gpurun_out/r2t_ncu.log 2>&1
tail -3 gpurun_out/r2t_ncu.log
ncu -i gpurun_out/r2t_group_full.ncu-rep --page raw --csv > gpurun_out/r2t_group_raw.csv 2>/dev/null
ncu -i gpurun_out/r2t_group_full.ncu-rep --page source --csv > gpurun_out/r2t_group_source.csv 2>/dev/null
ls -la gpurun_out/r2t_*
