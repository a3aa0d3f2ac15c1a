cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -3 gpurun_out/r2_bench_n$N.err)
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n$N.json 2> gpurun_out/r2_bench_ref_n$N.err; tail -2 gpurun_out/r2_bench_ref_n$N.err; cut -c1-300 gpurun_out/r2_bench_ref_n$N.json)
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
for k in ('value','n_gpus','ms_per_step','e2e','block_exit_ms','block_exit_alone','block_exit_host_breakdown_ms','cfg3_wq4','cfg5_70b_w4a16'):
    print(k, json.dumps(d.get(k))[:600])
print(json.dumps(d['extras']['w8a8_linear_8192x14336x4096'])[:500])
PY
