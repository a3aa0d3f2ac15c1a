#!/bin/bash
# round 2, final evidence on ONE GPU: full pytest, the default bench line, the reference arm, memcheck, the launch list of
# the bench command and a full ncu capture of one decoder layer (-> profiles/r02_ncu_full_layer.md, r02_traffic.json)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
(time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rs > gpurun_out/r3f_pytest.log 2>&1; tail -6 gpurun_out/r3f_pytest.log)
(time timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; tail -3 gpurun_out/r3f_bench.err)
(time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3f_bench_reference.json 2> gpurun_out/r3f_bench_reference.err; cut -c1-400 gpurun_out/r3f_bench_reference.json)
(time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py > gpurun_out/r3f_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r3f_sanitizer_memcheck.log)
(time timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/r3f_launches.csv python bench.py --steps 2 --warmup 1 --layers 2 --skip-extras --skip-cpu-baseline --skip-drop-in --skip-compiled-baseline > gpurun_out/r3f_launches_bench.log 2>&1; tail -2 gpurun_out/r3f_launches_bench.log | cut -c1-300)
if [ "${FULL_NCU:-0}" = 1 ]; then
(time timeout 900 ncu --set full --clock-control none -k regex:'ffq|calq|w8a8' -o gpurun_out/r3f_layer -f python tools/prof_layer.py 2 > gpurun_out/r3f_ncu_layer.log 2>&1; tail -3 gpurun_out/r3f_ncu_layer.log)
ncu -i gpurun_out/r3f_layer.ncu-rep --page raw --csv --print-units base > gpurun_out/r3f_layer_raw.csv 2>/dev/null
ls -la gpurun_out/r3f_layer.ncu-rep gpurun_out/r3f_layer_raw.csv
rm -f gpurun_out/r3f_layer.ncu-rep          # the report itself is over the 64 MiB that travel back; the raw page is what the summary reads
fi
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3f_bench.json').read().splitlines() if l.startswith('{')][-1])
for k in ('value','ms_per_step','e2e','ablation','block_exit_ms','cfg3_wq4','cfg5_70b_w4a16','kernels','roofline','gpu_launches_per_step'):
    print(k, json.dumps(d.get(k))[:600])
PY
