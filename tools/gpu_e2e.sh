# e2e check after an I/O change: GPU parity tests (host-buffer path) + bench line
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout -k 10 900 python bench.py --no-cpu-baseline 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json
tail -3 gpurun_out/bench_err.log
