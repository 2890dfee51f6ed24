# round 2: pass 1 with acq_rel tickets (no sc fences); full GPU test suite; probe; bench
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 python tools/probe.py --scene S3 --n 256 --steps 2 2>&1 | tail -12 | tee gpurun_out/r02_probe_s3_256_v4.log
timeout -k 10 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu_v4.log
timeout -k 10 900 python bench.py --no-cpu-baseline 2> gpurun_out/r02_bench_err_v4.log | tee gpurun_out/r02_bench_n1_v4.json
tail -3 gpurun_out/r02_bench_err_v4.log
