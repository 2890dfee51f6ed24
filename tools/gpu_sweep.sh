set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for m in "6 6" "6 7" "6 8" "7 6" "8 6" "8 8"; do
  set -- $m
  echo "=== PS_PASS1_OCC=$1 PS_PASS2_OCC=$2" | tee -a gpurun_out/sweep_occ.log
  PS_PASS1_OCC=$1 PS_PASS2_OCC=$2 timeout -k 10 300 python tools/probe.py --scene S3 --n 256 --steps 2 --reps 30 2>&1 | grep -E "pass1|pass2|cg_iter" | tail -6 | tee -a gpurun_out/sweep_occ.log
done
