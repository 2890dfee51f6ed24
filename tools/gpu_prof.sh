# ncu evidence for profiles/: launch list of one step + full captures of the hot kernels (single GPU)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_step.csv \
    python tools/probe.py --scene S3 --n 256 --steps 1 --reps 1 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k 'regex:pass[12]_kernel|cg_update_xr_kernel|cg_update_p_kernel|reduced_moments_kernel|reduced_expand_kernel' -s 24 -c 6 \
    -f -o gpurun_out/prof_hot python tools/probe.py --scene S3 --n 256 --steps 1 --reps 1 > gpurun_out/ncu_hot.log 2>&1
tail -3 gpurun_out/ncu_hot.log
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k 'regex:gram_moments_kernel|region_factor_kernel' -c 2 \
    -f -o gpurun_out/prof_gram python tools/probe.py --scene S3 --n 256 --steps 1 --reps 1 > gpurun_out/ncu_gram.log 2>&1
ls -la gpurun_out
