# round 2, 1 GPU, final state: full GPU suite, smoke(), ncu --set full of the four kernels of a CG iteration, the bench line
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_v14.log
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke_v14.log
timeout -k 10 200 ncu --set full --clock-control none --import-source on -k regex:"pass1_kernel|reduced_region_kernel|pass2_kernel|cg_update_kernel" --launch-skip 80 --launch-count 4 -o gpurun_out/r02_ncu_hot_v14 -f python tools/probe.py --scene S3 --n 256 --steps 1 --reps 1 > gpurun_out/r02_ncu_hot_v14.log 2>&1
tail -2 gpurun_out/r02_ncu_hot_v14.log
timeout -k 10 300 python bench.py 2> gpurun_out/r02_bench_err_v14.log | tee gpurun_out/r02_bench_n1_v14.json | cut -c1-300
tail -2 gpurun_out/r02_bench_err_v14.log
