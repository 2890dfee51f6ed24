# N-GPU check: distributed parity tests, then bench at 1..N ranks (usage: bash tools/gpu_multi.sh N [skiptests])
set -x
N=${1:-2}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv
if [ -z "$2" ]; then
timeout -k 10 900 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_dist_n$N.log
fi
timeout -k 10 600 python tools/probe.py --scene S3 --n 256 --steps 2 2>&1 | tail -9 | tee gpurun_out/probe_s3_256.log
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_err_n1.log | tee gpurun_out/bench_n1_multi.json
tail -3 gpurun_out/bench_err_n1.log
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
        bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/bench_err_n$n.log | tee gpurun_out/bench_n$n.json
    tail -5 gpurun_out/bench_err_n$n.log
    PS_COMM=nccl timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
        bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/bench_err_nccl_n$n.log | tee gpurun_out/bench_nccl_n$n.json
  fi
done
ls -la gpurun_out
