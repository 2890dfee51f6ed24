# round 2, 1 GPU (last minutes): one-CTA-per-region kernel with the regions launched longest first (PS_REGION_LPT), A/B + the GPU suite with it on
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout -k 5 60 python tools/probe_region.py; PS_REGION_LPT=1 timeout -k 5 60 python tools/probe_region.py; PS_REGION_LPT=1 PS_REGION_VARIANT=1 timeout -k 5 60 python tools/probe_region.py ) 2>&1 | grep "LPT=" | tee gpurun_out/r02_probe_region_lpt_v15.log
PS_REGION_LPT=1 timeout -k 10 230 python -m pytest tests -q -x -m gpu 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu_v15_lpt.log
