#!/usr/bin/env bash
# GPU test suite + C2 bench line (+ optional extra command): bash tools/gpu_tests.sh <tag> [pytest args]
set -u
tag="${1:-rX}"; shift || true
out=gpurun_out
mkdir -p "$out"
timeout 1200 python -m pytest tests -m gpu -x -q "$@" > "$out/${tag}_pytest_gpu.log" 2>&1
echo "pytest rc $?"
tail -15 "$out/${tag}_pytest_gpu.log"
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > "$out/${tag}_bench_c2.json" 2> "$out/${tag}_bench_c2.err"
echo "bench rc $?"
python - "$out/${tag}_bench_c2.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("ms/step", d["ms_per_step"], "e2e", d.get("e2e"))
for k in d.get("kernels", []):
    print("   ", {a: (round(b, 4) if isinstance(b, float) else b) for a, b in k.items()})
PY
