# round 2, first GPU check of the fused hot path: parity tests, probe, short bench, launch list
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc; free -g | head -2
timeout -k 10 1500 python -m pytest tests -q -x -m gpu --durations=8 2>&1 | tail -30 | tee gpurun_out/r02_pytest_gpu_v1.log
timeout -k 10 600 python tools/probe.py --scene S3 --n 256 --steps 2 2>&1 | tail -14 | tee gpurun_out/r02_probe_s3_256_v1.log
PS_PASS1_STATIC=1 timeout -k 10 600 python tools/probe.py --scene S3 --n 256 --steps 2 2>&1 | tail -14 | tee gpurun_out/r02_probe_s3_256_v1_static.log
timeout -k 10 900 python bench.py --no-cpu-baseline 2> gpurun_out/r02_bench_err_v1.log | tee gpurun_out/r02_bench_n1_v1.json
tail -3 gpurun_out/r02_bench_err_v1.log
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_step_v1.csv \
    python tools/probe.py --scene S3 --n 256 --steps 1 --reps 1 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
