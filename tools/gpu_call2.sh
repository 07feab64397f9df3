#!/usr/bin/env bash
# ncu --set full of the three Jacobi shapes through the torch-free harness (seconds each)
set -u
tag="${1:-r2b}"
out=gpurun_out
mkdir -p "$out"
build/jacobi_check /dev/null 1024 196 196 && \
ncu --set full --clock-control none --import-source on -k regex:jacobi_rows -c 1 -f -o "$out/${tag}_full_jacobi_proc" build/jacobi_check /dev/null 1024 196 196 > "$out/${tag}_ncu_proc.log" 2>&1
echo "proc rc $?"
build/jacobi_check /dev/null 16 384 384 && \
ncu --set full --clock-control none --import-source on -k regex:jacobi_rows -c 1 -f -o "$out/${tag}_full_jacobi_eig" build/jacobi_check /dev/null 16 384 384 > "$out/${tag}_ncu_eig.log" 2>&1
echo "eig rc $?"
build/jacobi_check /dev/null 48 384 384 174 && \
ncu --set full --clock-control none --import-source on -k regex:jacobi_rows -c 1 -f -o "$out/${tag}_full_jacobi_kxk" build/jacobi_check /dev/null 48 384 384 174 > "$out/${tag}_ncu_kxk.log" 2>&1
echo "kxk rc $?"
build/jacobi_check chol 1024 196 384 && \
ncu --set full --clock-control none --import-source on -k regex:cholesky -c 1 -f -o "$out/${tag}_full_chol" build/jacobi_check chol 1024 196 384 > "$out/${tag}_ncu_chol.log" 2>&1
echo "chol rc $?"
ls -la $out/*.ncu-rep
