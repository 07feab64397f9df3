#!/usr/bin/env bash
# Token-Gram kernels: their tests, then an ncu launch list of the statistics stage of one C2 step.
#   bash tools/gpu_gram.sh <tag>
set -u
tag="${1:-rX}"
out=gpurun_out
mkdir -p "$out"
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gram or merge or colsum" 2>&1 | tail -5
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  -k regex:'gram|colsum|mu_tile|rough_mean' --log-file "$out/${tag}_gram_launches.csv" \
  python tools/profile_step.py c2 256 spectral > /dev/null 2>&1
python - "$out/${tag}_gram_launches.csv" <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
head = rows[0]
ki, gi, vi = head.index("Kernel Name"), head.index("Grid Size"), head.index("Metric Value")
ui = head.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    k = (r[ki].split("(")[0], r[gi])
    a = acc.setdefault(k, [0.0, 0])
    a[0] += v; a[1] += 1
for k, (t, n) in acc.items():
    print(k, f"{t:.1f} us total {n} launches")
PY
