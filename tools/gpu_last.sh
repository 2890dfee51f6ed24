cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 270 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^Starting\|^Done\|^$\|threads\|^CG \|^Solve" | tail -8 > gpurun_out/pytest_gpu_last.log; cat gpurun_out/pytest_gpu_last.log
