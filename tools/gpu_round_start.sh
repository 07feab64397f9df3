#!/usr/bin/env bash
# One gpurun call that banks the round's baseline evidence before any kernel work:
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_start.sh r2'
# Order = value per GPU-minute: the torch-free kernel A/Bs (seconds), bench line, ncu launch list, full captures,
# the GPU test suite, then the remaining opt-in experiment that needs torch.
# On a fresh box the first `import torch` + eager module loading + backbone features take ~2 min,
# so no step gets less than 300 s (round 1 lost its last launch-list refresh to a 110 s limit).
set -u
tag="${1:-rN}"
out=gpurun_out
mkdir -p "$out"

echo "== Jacobi split experiment (torch-free harness, seconds): default vs 2 / 4 CTAs per problem"
if [ -x build/jacobi_check ]; then
  timeout 20 build/jacobi_check /tmp/j0.bin > "$out/${tag}_jacobi_split.log" 2>&1
  BASD_JACOBI_SPLIT=2 timeout 20 build/jacobi_check /tmp/j2.bin >> "$out/${tag}_jacobi_split.log" 2>&1
  BASD_JACOBI_SPLIT=4 timeout 20 build/jacobi_check /tmp/j4.bin >> "$out/${tag}_jacobi_split.log" 2>&1
  cmp /tmp/j0.bin /tmp/j2.bin >> "$out/${tag}_jacobi_split.log" 2>&1 && echo "split 2 bitwise equal" >> "$out/${tag}_jacobi_split.log"
  cmp /tmp/j0.bin /tmp/j4.bin >> "$out/${tag}_jacobi_split.log" 2>&1 && echo "split 4 bitwise equal" >> "$out/${tag}_jacobi_split.log"
  rm -f /tmp/j0.bin /tmp/j2.bin /tmp/j4.bin
  # the k x k principal-angle SVD shape: 48 problems, 384-wide allocation, active size 174
  for sp in "" 2 4; do BASD_JACOBI_SPLIT=$sp timeout 20 build/jacobi_check /dev/null 48 384 384 174; done \
    >> "$out/${tag}_jacobi_split.log" 2>&1
  # the rank-aware entry point (C3: 48 non-zero rows of 196)
  for sp in "" 2 4; do BASD_JACOBI_SPLIT=$sp timeout 20 build/jacobi_check /dev/null 1024 196 196 0 48; done \
    >> "$out/${tag}_jacobi_split.log" 2>&1
  cat "$out/${tag}_jacobi_split.log"
  echo "== register-resident pivoted Cholesky experiment: default vs BASD_CHOL_REG=1|2 (full rank, rank 48, n = 150)"
  for args in "1024 196 384" "1024 196 48" "256 150 300"; do
    timeout 30 build/jacobi_check chol $args
    BASD_CHOL_REG=1 timeout 30 build/jacobi_check chol $args      # four lanes per row
    BASD_CHOL_REG=2 timeout 30 build/jacobi_check chol $args      # two lanes per row
  done > "$out/${tag}_chol_reg.log" 2>&1
  cat "$out/${tag}_chol_reg.log"
fi

echo "== bench c2"
timeout 400 python bench.py --steps 5 --warmup 3 > "$out/${tag}_bench_c2.json" 2> "$out/${tag}_bench_c2.err"
echo "bench rc $?"

echo "== profiled program without ncu"
if timeout 400 python tools/profile_step.py c2 256 > "$out/${tag}_profile_step.log" 2>&1; then
  echo "== ncu launch list"
  timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file "$out/${tag}_launches_c2_b256.csv" python tools/profile_step.py c2 256 > "$out/${tag}_ncu_launches.log" 2>&1
  echo "launch list rc $?"
  echo "== ncu --set full, top kernels, batch 64"
  timeout 300 python tools/profile_step.py c2 64 > "$out/${tag}_profile_step_b64.log" 2>&1 && \
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"oe8|gemm_tc3|pivoted_cholesky_left4|token_gram_tc|mix_interp|weight_grad_kernel" -c 24 \
    -o "$out/${tag}_full_c2_b64" python tools/profile_step.py c2 64 > "$out/${tag}_ncu_full.log" 2>&1
  echo "full capture rc $?"
  if [ -f "$out/${tag}_full_c2_b64.ncu-rep" ]; then
    ncu -i "$out/${tag}_full_c2_b64.ncu-rep" --page raw --csv > "$out/${tag}_full_c2_b64_raw.csv" 2>/dev/null
  fi
else
  echo "profile_step failed or timed out: see $out/${tag}_profile_step.log"
fi

echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > "$out/${tag}_pytest_gpu.log" 2>&1
echo "pytest rc $?"
tail -5 "$out/${tag}_pytest_gpu.log"

echo "== one-pass weight-gradient experiment (BASD_WGRAD_ONEPASS=1): kernel test, then its time in the bench timeline"
BASD_WGRAD_ONEPASS=1 timeout 400 python -m pytest tests/test_kernels_gpu.py -q -k "mix_interp_and_weight_grad" \
  > "$out/${tag}_wgrad_onepass_test.log" 2>&1
echo "one-pass kernel test rc $?"; tail -2 "$out/${tag}_wgrad_onepass_test.log"
BASD_WGRAD_ONEPASS=1 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline \
  > "$out/${tag}_bench_c2_wgrad_onepass.json" 2> "$out/${tag}_bench_c2_wgrad_onepass.err"
python - "$out/${tag}_bench_c2.json" "$out/${tag}_bench_c2_wgrad_onepass.json" <<'PY'
import json, sys
for path in sys.argv[1:]:
    try:
        d = json.load(open(path))
        wg = [k for k in d.get("kernels", []) if k["kernel"] == "basd_weight_grad"]
        print(path, "ms/step", d["ms_per_step"], "weight_grad ms", wg[0]["ms"] if wg else None)
    except Exception as e:
        print(path, "unreadable:", e)
PY
