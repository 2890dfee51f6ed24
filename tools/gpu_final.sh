# final single-GPU check of the round: the whole GPU suite (incl. the compiled-reference parity tests), smoke, bench (both arms)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ls -la oracle/_ref
timeout -k 10 1500 python -m pytest tests -q -m gpu --durations=8 2>&1 | grep -v "^Starting\|^Done\|^$\|threads\|^CG \|^Solve" | tail -30 > gpurun_out/pytest_gpu_final.log; tail -14 gpurun_out/pytest_gpu_final.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_final.log
timeout -k 10 900 python bench.py 2> gpurun_out/bench_err_final.log | tee gpurun_out/bench_n1_final.json | cut -c1-300
tail -3 gpurun_out/bench_err_final.log
timeout -k 10 600 python bench.py --impl reference --steps 1 --warmup 0 2> gpurun_out/bench_ref_err_final.log | tee gpurun_out/bench_ref_n1_final.json | cut -c1-300
