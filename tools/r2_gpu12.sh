# round 2, 1 GPU: Gram moments on DMMA (parity through the full GPU suite, stage time against the scalar kernel, ncu of the kernel), S3 bench with the final
# defaults, configs[3] S4 384^3 and configs[4] S5 512x256x256 with the tile sweep 8 / 16 / 32 on one GPU
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PS_GRAM_DMMA=0 timeout -k 10 300 python tools/probe.py --scene S3 --n 256 --steps 2 --reps 5 2>&1 | grep -E "stages|step 1" | tee gpurun_out/r02_probe_gram_scalar_v12.log
timeout -k 10 300 python tools/probe.py --scene S3 --n 256 --steps 2 --reps 30 2>&1 | tail -12 | tee gpurun_out/r02_probe_s3_256_v12.log
timeout -k 10 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu_v12.log
timeout -k 10 600 python bench.py --no-cpu-baseline 2> gpurun_out/r02_bench_err_v12.log | tee gpurun_out/r02_bench_n1_v12.json
tail -2 gpurun_out/r02_bench_err_v12.log
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gram_moments -c 1 -o gpurun_out/r02_ncu_gram_dmma_v12 -f python tools/probe.py --scene S3 --n 256 --steps 1 --reps 1 > gpurun_out/r02_ncu_gram_v12.log 2>&1
PS_GRAM_DMMA=0 timeout -k 10 300 ncu --set full --clock-control none -k regex:gram_moments -c 1 -o gpurun_out/r02_ncu_gram_scalar_v12 -f python tools/probe.py --scene S3 --n 256 --steps 1 --reps 1 >> gpurun_out/r02_ncu_gram_v12.log 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_step_v12.csv python tools/probe.py --scene S3 --n 256 --steps 1 --reps 1 > gpurun_out/r02_ncu_launch_v12.log 2>&1
timeout -k 10 900 python bench.py --scene S4 --steps 2 --warmup 1 --kernel-reps 10 --no-cpu-baseline 2> gpurun_out/r02_bench_err_S4_n1_v12.log | tee gpurun_out/r02_bench_S4_n1_v12.json
tail -2 gpurun_out/r02_bench_err_S4_n1_v12.log
for t in 16 8 32; do
timeout -k 10 900 python bench.py --scene S5 --tile $t --steps 2 --warmup 1 --kernel-reps 10 --no-cpu-baseline 2> gpurun_out/r02_bench_err_S5_t${t}_n1_v12.log | tee gpurun_out/r02_bench_S5_t${t}_n1_v12.json
tail -2 gpurun_out/r02_bench_err_S5_t${t}_n1_v12.log
done
