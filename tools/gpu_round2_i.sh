set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
(timeout 120 python tools/prof_gemm_roles.py > gpurun_out/r2i_gemm_roles.jsonl 2>&1; cat gpurun_out/r2i_gemm_roles.jsonl | cut -c1-1200)
(FFQ_GEMM_DIRECT_STORE=1 timeout 120 python tools/prof_gemm_roles.py 8192 14336 4096 2>&1 | cut -c1-1200)
(timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -rs > gpurun_out/r2i_pytest_2gpu.log 2>&1; tail -6 gpurun_out/r2i_pytest_2gpu.log)
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --skip-compiled-baseline --skip-cpu-baseline > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; tail -5 gpurun_out/r2i_bench_n2.err)
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench_n2.json').read().strip().splitlines()[-1])
for k in ('value','n_gpus','ms_per_step','e2e','block_exit_ms','block_exit_alone','block_exit_host_breakdown_ms'):
    print(k, json.dumps(d.get(k))[:700])
PY
