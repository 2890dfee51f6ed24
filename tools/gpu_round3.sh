# single-GPU round check: all GPU tests (incl. the reference-code solve stage), smoke, bench (both arms), ncu launch list of bench.py
# itself, full captures of the five kernels of the CG iteration
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
ls -la oracle/_ref
timeout -k 10 1200 python -m pytest tests -q -m gpu --durations=6 2>&1 | tail -24 | tee gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout -k 10 900 python bench.py 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json
tail -3 gpurun_out/bench_err.log
timeout -k 10 600 python bench.py --impl reference --steps 1 --warmup 0 2> gpurun_out/bench_ref_err.log | tee gpurun_out/bench_ref_n1.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --kernel-reps 2 > gpurun_out/ncu_launch_bench.log 2>&1
tail -2 gpurun_out/ncu_launch_bench.log | cut -c1-300
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k 'regex:pass[12]_kernel|cg_update_r_kernel|cg_update_xp_kernel|reduced_region_kernel' -s 40 -c 5 \
    -f -o gpurun_out/prof_hot python tools/probe.py --scene S3 --n 256 --steps 1 --reps 1 > gpurun_out/ncu_hot.log 2>&1
tail -3 gpurun_out/ncu_hot.log
ls -la gpurun_out
