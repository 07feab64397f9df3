#!/usr/bin/env bash
# Final single-GPU lines of the round: C2 (headline, with the CPU arm), C1, C3, C4, the reference arm.
set -u
tag="${1:-rX}"
out=gpurun_out
mkdir -p "$out"
timeout 900 python bench.py --steps 20 --warmup 5 > "$out/${tag}_bench_c2.json" 2> "$out/${tag}_bench_c2.err"; echo "c2 rc $?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > "$out/${tag}_bench_c2_reference.json" 2> "$out/${tag}_bench_c2_reference.err"; echo "ref rc $?"
timeout 300 python bench.py --workload c1 --batch 16 --steps 50 --warmup 10 --no-cpu-baseline > "$out/${tag}_bench_c1.json" 2>/dev/null; echo "c1 rc $?"
timeout 300 python bench.py --workload c1 --batch 16 --steps 50 --warmup 10 --no-cpu-baseline --graph > "$out/${tag}_bench_c1_graph.json" 2>/dev/null; echo "c1 graph rc $?"
timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > "$out/${tag}_bench_c3.json" 2>/dev/null; echo "c3 rc $?"
timeout 900 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > "$out/${tag}_bench_c4.json" 2>/dev/null; echo "c4 rc $?"
python - "$out" "$tag" <<'PY'
import json, sys, glob
out, tag = sys.argv[1:3]
for p in sorted(glob.glob(f"{out}/{tag}_bench_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        print(p, "unreadable", e); continue
    print(p.split("/")[-1], "value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"),
          "cpu", (d.get("cpu_baseline") or {}).get("value"), "clocks", d.get("clocks"))
PY
