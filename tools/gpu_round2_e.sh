set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2e_pytest.log 2>&1; tail -8 gpurun_out/r2e_pytest.log)
cd tools
(timeout 200 python bench_group.py; FFQ_CALQ_GROUP_SUBWARP=1 timeout 200 python bench_group.py) > ../gpurun_out/r2e_group.log 2>&1
(echo FAT; timeout 200 python bench_calq.py; echo NOFAT; FFQ_CALQ_TENSOR_FAT=0 timeout 200 python bench_calq.py) > ../gpurun_out/r2e_calq.log 2>&1
cd ..
cat gpurun_out/r2e_group.log gpurun_out/r2e_calq.log
