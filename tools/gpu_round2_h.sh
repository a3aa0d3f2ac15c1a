set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
(timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -rs > gpurun_out/r2h_pytest_2gpu.log 2>&1; tail -12 gpurun_out/r2h_pytest_2gpu.log)
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --skip-compiled-baseline > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; tail -5 gpurun_out/r2h_bench_n2.err)
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench_n2.json').read().strip().splitlines()[-1])
for k in ('value','n_gpus','ms_per_step','e2e','ablation','block_exit_ms','cfg3_wq4','cfg5_70b_w4a16'):
    print(k, json.dumps(d.get(k))[:600])
print(json.dumps(d['extras']['w8a8_linear_8192x14336x4096'])[:500])
PY
