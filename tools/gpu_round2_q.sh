cd $GRAFT_REPO_ROOT
timeout 100 python tools/prof_calq_phases.py 2>&1 | cut -c1-900
(cd tools; timeout 100 python bench_calq.py 2>&1 | tail -9)
(timeout 600 python bench.py --steps 10 --warmup 3 --skip-extras --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','block_exit_ms','block_exit_alone','block_exit_host_breakdown_ms','kernels','gpu_launches_per_step'):
    print(k, json.dumps(d.get(k))[:500])
")
