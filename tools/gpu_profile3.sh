#!/usr/bin/env bash
# Final-code refresh of the round's ncu evidence: launch list of one C2 B=256 step (backbone features) and
# `--set full` of every launch of the kernels that changed after tools/gpu_profile.sh / gpu_profile2.sh ran
# (3xTF32 GEMM, token Gram + its reduce kernels, fp64 rotation, cluster Jacobi).  bash tools/gpu_profile3.sh <tag>
set -u
tag="${1:-rX}"
out=gpurun_out
mkdir -p "$out"
timeout 400 python tools/profile_step.py c2 256 > "$out/${tag}_profile_step.log" 2>&1 || { echo "profile_step failed"; tail -5 "$out/${tag}_profile_step.log"; exit 1; }
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file "$out/${tag}_launches_c2_b256.csv" python tools/profile_step.py c2 256 > "$out/${tag}_ncu_launches.log" 2>&1
echo "launch list rc $?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none \
  -k "regex:gemm_tc3_tma|token_gram_tc|gram_reduce|colsum_fold|dgemm_kernel|jacobi_rows_oe8_cluster" -c 80 -f \
  -o /tmp/${tag}_final python tools/profile_step.py c2 256 > "$out/${tag}_ncu_final.log" 2>&1
echo "full rc $?"
[ -f /tmp/${tag}_final.ncu-rep ] && ncu -i /tmp/${tag}_final.ncu-rep --page raw --csv > "$out/${tag}_full_final_raw.csv" 2>/dev/null
ls -la "$out"/${tag}_full_final_raw.csv "$out"/${tag}_launches_c2_b256.csv
du -sh "$out"
