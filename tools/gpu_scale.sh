# N-GPU scaling check: distributed parity tests, then bench.py at 2..N ranks over peer memory (usage: bash tools/gpu_scale.sh N [skiptests])
set -x
N=${1:-4}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv
if [ -z "$2" ]; then
timeout -k 10 900 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_dist_n$N.log
fi
for n in 2 4 8; do
  if [ $n -le $N ]; then
    PS_TRACE=60 timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
        bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/bench_err_n$n.log | tee gpurun_out/bench_n$n.json
    grep "ps trace rank 0" gpurun_out/bench_err_n$n.log | tail -2
    tail -3 gpurun_out/bench_err_n$n.log
  fi
done
