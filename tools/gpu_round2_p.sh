set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
(time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rs > gpurun_out/r2p_pytest.log 2>&1; tail -6 gpurun_out/r2p_pytest.log)
(time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -3 gpurun_out/r2p_bench.err)
(time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p_bench_reference.json 2> gpurun_out/r2p_bench_reference.err; tail -3 gpurun_out/r2p_bench_reference.err; cut -c1-600 gpurun_out/r2p_bench_reference.json)
(time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py > gpurun_out/r2p_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2p_sanitizer_memcheck.log)
(time timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 1 --layers 2 --skip-extras --skip-cpu-baseline > gpurun_out/r2p_launches_bench.log 2>&1; tail -2 gpurun_out/r2p_launches_bench.log | cut -c1-300)
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','ablation','block_exit_ms','block_exit_alone','block_exit_host_breakdown_ms','cfg3_wq4','cfg5_70b_w4a16','kernels','roofline'):
    print(k, json.dumps(d.get(k))[:700])
PY
