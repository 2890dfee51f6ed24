# GPU check of the reference-code parity tests (classifier, matrix blocks, solve stage, export) + the whole GPU suite
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ls -la oracle/_ref
timeout -k 10 1500 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -40 > gpurun_out/pytest_gpu_r4.log; tail -16 gpurun_out/pytest_gpu_r4.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_r4.log
