#!/usr/bin/env bash
# Targeted `--set full` captures of the kernels the whole-step capture did not reach inside its time limit.
set -u
tag="${1:-rX}"
out=gpurun_out
mkdir -p "$out"
J=build/jacobi_check
cap() {  # name, regex, count, command...
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" -c "$cnt" -f -o /tmp/${tag}_$name "$@" > "$out/${tag}_ncu_$name.log" 2>&1
  echo "$name rc $?"
  if [ -f /tmp/${tag}_$name.ncu-rep ]; then
    ncu -i /tmp/${tag}_$name.ncu-rep --page raw --csv > "$out/${tag}_full_${name}_raw.csv" 2>/dev/null
  fi
}
$J /dev/null 1024 196 196 && cap jacobi_proc jacobi_rows 1 $J /dev/null 1024 196 196
[ -f /tmp/${tag}_jacobi_proc.ncu-rep ] && ncu -i /tmp/${tag}_jacobi_proc.ncu-rep --page source --csv > "$out/${tag}_source_jacobi_proc.csv" 2>/dev/null
cap jacobi_kxk jacobi_rows 1 $J /dev/null 48 384 384 174
cap chol_reg cholesky 1 $J chol 1024 196 384
timeout 300 python tools/profile_step.py c2 256 spectral > "$out/${tag}_profile_step_spectral.log" 2>&1 && \
timeout 900 ncu --profile-from-start off --set full --clock-control none \
  -k "regex:mix_interp|weight_grad_onepass|weight_grad_finish|weighted_center|gemm_tc3|grad_prep|rows_finish|mix_rows|extract_diag" -c 45 -f \
  -o /tmp/${tag}_stepk python tools/profile_step.py c2 256 spectral > "$out/${tag}_ncu_stepk.log" 2>&1
echo "stepk rc $?"
[ -f /tmp/${tag}_stepk.ncu-rep ] && ncu -i /tmp/${tag}_stepk.ncu-rep --page raw --csv > "$out/${tag}_full_stepk_raw.csv" 2>/dev/null
ls -la "$out"/${tag}_full_*_raw.csv "$out"/${tag}_source_jacobi_proc.csv
du -sh "$out"
