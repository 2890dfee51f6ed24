# round 2, 2 GPUs, last seconds: a different scene on the same distributed handle (peer transport): arena layout follows the new system size, slab-local setup falls back to replicated
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 50 python -m pytest tests/test_gpu_distributed.py -q -x -m gpu -k "then_shrunk-2-peer" 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_dist_next_v16.log
