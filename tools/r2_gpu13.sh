# round 2, 8 GPUs (one shot): the metric's step at N = 8 and 4, distributed parity at 8 / 4 / 3 ranks, configs[3] S4 384^3 and configs[4] S5 512x256x256 at N = 8
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -k 10 150 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02_bench_err_n8_v13.log | tee gpurun_out/r02_bench_n8_v13.json | cut -c1-400
tail -2 gpurun_out/r02_bench_err_n8_v13.log
timeout -k 10 150 $TR --nproc-per-node 4 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02_bench_err_n4_v13.log | tee gpurun_out/r02_bench_n4_v13.json | cut -c1-400
timeout -k 10 200 python -m pytest tests/test_gpu_distributed.py -q -x -m gpu -k "s3_128-8-peer or s3_128-4-peer or blob64_tile16-4-peer or layers33-3-peer" 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu_dist_n8_v13.log
timeout -k 10 200 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --scene S4 --steps 2 --warmup 1 --kernel-reps 10 --no-cpu-baseline 2> gpurun_out/r02_bench_err_S4_n8_v13.log | tee gpurun_out/r02_bench_S4_n8_v13.json | cut -c1-400
tail -2 gpurun_out/r02_bench_err_S4_n8_v13.log
timeout -k 10 150 $TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --scene S5 --steps 2 --warmup 1 --kernel-reps 10 --no-cpu-baseline 2> gpurun_out/r02_bench_err_S5_n8_v13.log | tee gpurun_out/r02_bench_S5_n8_v13.json | cut -c1-400
PS_TRACE=100 timeout -k 10 100 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep "ps trace rank 3" | tail -2 | tee gpurun_out/r02_trace_n8_v13.log
timeout -k 10 100 python -m pytest tests/test_gpu_multi.py -q -x -m gpu -k "s3_128-8" 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_multi_n8_v13.log
