cd $GRAFT_REPO_ROOT
(timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; tail -2 gpurun_out/r2y_bench.err)
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','ablation','block_exit_ms','block_exit_alone','cfg3_wq4','cfg5_70b_w4a16','kernels','gpu_launches_per_step'):
    print(k, json.dumps(d.get(k))[:600])
for k,v in d['extras'].items(): print(k, json.dumps(v)[:260])
PY
