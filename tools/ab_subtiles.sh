#!/bin/bash
# A/B of the sub-tile interleave on the bench workload: tools/ab_subtiles.sh [ROWS] ["subtile list"]
ROWS=${1:-125000}; SUBS=${2:-"1 2"}
mkdir -p gpurun_out
for s in $SUBS; do
  timeout 300 python bench.py --rows $ROWS --steps 3 --warmup 2 --no-cpu --no-e2e --subtiles $s > gpurun_out/ab_$s.json 2> gpurun_out/ab_$s.err || tail -3 gpurun_out/ab_$s.err
  python - "$s" "$ROWS" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
    print("AB subtiles", sys.argv[1], "rows", sys.argv[2], "ms/step", round(d["ms_per_step"], 1), "users/s", round(d["value"]), "frac", round(d["roofline"]["frac"], 3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("AB subtiles", sys.argv[1], "failed", e)
PY
done
