set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_ref_solve.py -q -m gpu -k "bicgstab or S2_beam or export" 2>&1 | tail -60 | tee gpurun_out/pytest_gpu_retest.log
for i in 1 2 3; do timeout -k 10 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "bicgstab" 2>&1 | tail -3; done | tee -a gpurun_out/pytest_gpu_retest.log
