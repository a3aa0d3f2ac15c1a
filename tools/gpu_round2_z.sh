cd $GRAFT_REPO_ROOT
(timeout 300 python -m pytest tests/test_qlinear_gpu.py tests/test_round2_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3)
timeout 200 python tools/bench_extras.py 2>&1 | grep w8a8 | cut -c1-400
