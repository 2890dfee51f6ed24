# round 2: barrier-free pass 1 (warp-level region term + expand phase), unroll A/B, parity subset, bench
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu 2>&1 | tail -5 | tee gpurun_out/r02_pytest_gpu_v3.log
timeout -k 10 600 python tools/probe.py --scene S3 --n 256 --steps 2 2>&1 | tail -12 | tee gpurun_out/r02_probe_s3_256_v3.log
PS_PASS1_UNROLL=1 timeout -k 10 600 python tools/probe.py --scene S3 --n 256 --steps 2 2>&1 | tail -9 | tee gpurun_out/r02_probe_s3_256_v3_unroll1.log
timeout -k 10 900 python bench.py --no-cpu-baseline 2> gpurun_out/r02_bench_err_v3.log | tee gpurun_out/r02_bench_n1_v3.json
tail -3 gpurun_out/r02_bench_err_v3.log
