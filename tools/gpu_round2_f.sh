set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2f_pytest.log 2>&1; tail -8 gpurun_out/r2f_pytest.log)
cd tools
(echo MT-default; timeout 200 python bench_w4a16.py; echo MT=1; FFQ_W4A16_MT=1 timeout 200 python bench_w4a16.py) > ../gpurun_out/r2f_w4a16.log 2>&1
cd ..
cat gpurun_out/r2f_w4a16.log
(timeout 900 python bench.py --steps 10 --warmup 3 --skip-compiled-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -5 gpurun_out/r2f_bench.err)
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
for k in ('value','ms_per_step','ablation','block_exit_ms','cfg3_wq4','cfg5_70b_w4a16','kernels'):
    print(k, json.dumps(d.get(k))[:700])
for k,v in d['extras'].items(): print(k, json.dumps(v)[:300])
PY
