#!/usr/bin/env bash
# N-GPU checks: bash tools/gpu_multi.sh <tag> <N> [workload]
set -u
tag="${1:-rX}"; n="${2:-2}"; wl="${3:-c2}"
out=gpurun_out
mkdir -p "$out"
nvidia-smi -L
if [ "$n" = "2" ]; then
  timeout 600 python -m pytest tests/test_dp_gpu.py -q -s > "$out/${tag}_dp_test.log" 2>&1
  echo "dp test rc $?"; tail -8 "$out/${tag}_dp_test.log"
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus "$n" --steps 10 --warmup 3 --workload "$wl" > "$out/${tag}_bench_${wl}_n${n}.json" 2> "$out/${tag}_bench_${wl}_n${n}.err"
echo "bench rc $?"
tail -3 "$out/${tag}_bench_${wl}_n${n}.err"
python - "$out/${tag}_bench_${wl}_n${n}.json" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print("n_gpus", d["n_gpus"], "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"],
          "h2d GB/s/gpu", d["e2e"].get("h2d_gb_per_s_per_gpu"), "staging", d["e2e"].get("staging"))
    print("dp_parity", d.get("dp_parity"))
PY
