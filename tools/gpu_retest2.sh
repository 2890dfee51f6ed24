set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3; do timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_ref_solve.py -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu_retest2_$i.log; tail -3 gpurun_out/pytest_gpu_retest2_$i.log; done
