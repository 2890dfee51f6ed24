# hunt for the one-off failure of test_gpu_bicgstab_fallback inside the full suite: initcheck / memcheck on the small solver tests, then the
# whole tests/test_gpu_parity.py file twice with the complete failure output
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 500 compute-sanitizer --tool initcheck --track-unused-memory no --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "bicgstab or blob40-" 2>&1 | grep -v "^$" | tail -60 > gpurun_out/sanitize_initcheck.log
tail -25 gpurun_out/sanitize_initcheck.log
timeout -k 10 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "bicgstab" 2>&1 | grep -v "^$" | tail -30 > gpurun_out/sanitize_memcheck.log
tail -8 gpurun_out/sanitize_memcheck.log
for i in 1 2; do timeout -k 10 600 python -m pytest tests/test_golden.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -70 > gpurun_out/pytest_gpu_full_$i.log; tail -4 gpurun_out/pytest_gpu_full_$i.log; done
