#!/usr/bin/env bash
# Round profile refresh on the final code: launch list of one C2 B=256 step + ONE `--set full` capture of every
# launch of that step.  bash tools/gpu_profile.sh <tag>   (CSV outputs under gpurun_out/, summarised by
# tools/ncu_summary.py into profiles/; the .ncu-rep is exported to CSV on the box and deleted: gpurun copies
# back at most 64 MiB)
set -u
tag="${1:-rX}"
out=gpurun_out
mkdir -p "$out"
timeout 400 python tools/profile_step.py c2 256 > "$out/${tag}_profile_step.log" 2>&1 || { echo "profile_step failed"; tail -5 "$out/${tag}_profile_step.log"; exit 1; }
echo "== launch list"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file "$out/${tag}_launches_c2_b256.csv" python tools/profile_step.py c2 256 > "$out/${tag}_ncu_launches.log" 2>&1
echo "launch list rc $?"
echo "== full capture of the whole step"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -f \
  -o /tmp/${tag}_full python tools/profile_step.py c2 256 > "$out/${tag}_ncu_full.log" 2>&1
echo "full rc $?"
if [ -f /tmp/${tag}_full.ncu-rep ]; then
  ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > "$out/${tag}_full_raw.csv" 2>/dev/null
  # per-instruction stall samples of the dominant kernel only (source page of everything is too large)
  ncu -i /tmp/${tag}_full.ncu-rep --page source --csv -k regex:jacobi_rows_oe8_kernel 2>/dev/null | head -6000 > "$out/${tag}_source_jacobi_oe8.csv"
  ls -la /tmp/${tag}_full.ncu-rep
fi
ls -la "$out"/${tag}_full_raw.csv "$out"/${tag}_launches_c2_b256.csv
du -sh "$out"
