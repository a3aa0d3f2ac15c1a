cd $GRAFT_REPO_ROOT
timeout 100 python tools/prof_calq_phases.py 2>&1 | cut -c1-900
(cd tools; timeout 100 python bench_calq.py 2>&1 | tail -9)
(timeout 600 python -m pytest tests/test_calibrate_gpu.py tests/test_round2_gpu.py tests/test_api_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3)
