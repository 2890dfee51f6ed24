cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 40 python -m pytest tests/test_set_params.py -q -x -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_reuse_v17.log
