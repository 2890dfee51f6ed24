set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" | tee gpurun_out/lscpu.txt
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json
tail -5 gpurun_out/bench_err.log
# launch list of one full step (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_step.csv \
    python tools/probe.py --scene S3 --n 256 --steps 1 --reps 1 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
# full capture of the hot kernels of one CG iteration
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:pass[12]_kernel|cg_update_xr_kernel|cg_update_p_kernel|moments_partial_kernel' -s 20 -c 5 \
    -f -o gpurun_out/prof_hot python tools/probe.py --scene S3 --n 256 --steps 1 --reps 1 > gpurun_out/ncu_hot.log 2>&1
tail -3 gpurun_out/ncu_hot.log
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:gram_partial_kernel' -c 1 \
    -f -o gpurun_out/prof_gram python tools/probe.py --scene S3 --n 256 --steps 1 --reps 1 > gpurun_out/ncu_gram.log 2>&1
ls -la gpurun_out
