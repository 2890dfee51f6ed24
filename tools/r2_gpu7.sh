# round 2, 2 GPUs: the ps_create_multi handle after the deferred-free fix (each case in a child process under a time-out, watchdog at 45 s),
# the adaptor on a 2-device handle, the distributed suite
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout -k 10 420 python -m pytest tests/test_gpu_multi.py tests/test_zz_adaptor.py -q -x -m gpu 2>&1 | tail -25 | tee gpurun_out/r02_pytest_gpu_multi_n2_v7.log
timeout -k 10 420 python -m pytest tests/test_gpu_distributed.py -q -x -m gpu 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu_dist_n2_v7.log
