set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:pass[12]_kernel|reduced_pass1_kernel|reduced_finish_kernel|reduced_expand_kernel' -s 20 -c 5 \
    -f -o gpurun_out/prof_hot2 python tools/probe.py --scene S3 --n 256 --steps 1 --reps 1 > gpurun_out/ncu_hot2.log 2>&1
tail -3 gpurun_out/ncu_hot2.log
