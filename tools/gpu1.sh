set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 600 python tools/probe.py --scene S3 --n 128 --steps 2 2>&1 | tail -30 | tee gpurun_out/probe_s3_128.log
timeout 900 python tools/probe.py --scene S3 --n 256 --steps 2 2>&1 | tail -30 | tee gpurun_out/probe_s3_256.log
