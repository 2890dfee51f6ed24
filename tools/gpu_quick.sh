# quick single-GPU iteration: parity tests (optionally a -k filter as $1) + probe (kernel rooflines)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -x -q -m gpu --durations=8 2>&1 | tail -16 | tee gpurun_out/pytest_gpu_quick.log
timeout -k 10 600 python tools/probe.py --scene S3 --n 256 --steps 2 2>&1 | tail -12 | tee gpurun_out/probe_s3_256.log
