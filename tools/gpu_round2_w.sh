cd $GRAFT_REPO_ROOT
for lib in build/ab/libffq_prev.so fastforward_b200/lib/libffq_b200.so; do
echo "== $lib"
(cd tools; FFQ_LIB_PATH=$GRAFT_REPO_ROOT/$lib timeout 100 python bench_calq.py 2>&1 | tail -9 | grep "sym=False" | cut -c1-150)
done
timeout 100 python tools/prof_calq_phases.py 2>&1 | cut -c1-900
(timeout 600 python -m pytest tests/test_calibrate_gpu.py tests/test_round2_gpu.py tests/test_api_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3)
