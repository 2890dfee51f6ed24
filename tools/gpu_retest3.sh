cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do timeout -k 10 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_gpu_bicgstab_fallback" 2>&1 | grep "^E  \|passed\|failed" | head -6; done 2>&1 | tee gpurun_out/bicgstab_flaky.log
