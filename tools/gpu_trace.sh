set -x
N=${1:-2}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PS_TRACE=60 timeout -k 10 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --kernel-reps 5 2>&1 | grep "ps trace" | tee gpurun_out/trace_n1.log
PS_TRACE=60 timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 1 --warmup 1 --kernel-reps 5 2>&1 | grep "ps trace" | tee gpurun_out/trace_peer_n$N.log
PS_COMM=nccl PS_TRACE=60 timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1 --warmup 1 --kernel-reps 5 2>&1 | grep "ps trace" | tee gpurun_out/trace_nccl_n$N.log
