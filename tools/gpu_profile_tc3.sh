#!/usr/bin/env bash
# `--set full` of the tc3 GEMM launches of one C2 step (raw page of all, source page of the longest).
set -u
tag="${1:-rX}"
out=gpurun_out
mkdir -p "$out"
timeout 300 python tools/profile_step.py c2 256 spectral > "$out/${tag}_profile_step.log" 2>&1 || { tail -5 "$out/${tag}_profile_step.log"; exit 1; }
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k "regex:gemm_tc3" -c 30 -f -o /tmp/${tag}_tc3 python tools/profile_step.py c2 256 spectral > "$out/${tag}_ncu_tc3.log" 2>&1
echo "ncu rc $?"
ncu -i /tmp/${tag}_tc3.ncu-rep --page raw --csv > "$out/${tag}_full_tc3_raw.csv" 2>/dev/null
python - "$out/${tag}_full_tc3_raw.csv" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
def col(name): return hdr.index(name)
ids = [int(r[col("ID")]) for r in rows[2:]]
keys = ["launch__grid_size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
best = None
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d["ID"], d.get("Grid Size"), " ".join(f"{k.split('.')[0][-22:]}={d.get(k)}" for k in keys[1:]))
    t = float(d["gpu__time_duration.sum"].replace(",", ""))
    if best is None or t > best[0]: best = (t, d["ID"])
print("longest id", best)
open(sys.argv[1] + ".longest", "w").write(best[1])
PY
id=$(cat "$out/${tag}_full_tc3_raw.csv.longest")
ncu -i /tmp/${tag}_tc3.ncu-rep --page source --csv --launch-skip "$id" --launch-count 1 > "$out/${tag}_source_tc3.csv" 2>/dev/null
ls -la "$out"/${tag}_*tc3*
