#!/bin/bash
# usage: tools/quick_bench.sh ROWS "cluster list" [workload]  -> one compact line per run
ROWS=${1:-37888}; CL=${2:-"1 2"}; WL=${3:-cfg5}
mkdir -p gpurun_out
for c in $CL; do
  python bench.py --workload $WL --rows $ROWS --steps 2 --warmup 1 --no-cpu --no-e2e --cluster $c > gpurun_out/qb_$c.json 2> gpurun_out/qb_$c.err || tail -3 gpurun_out/qb_$c.err
  python - "$c" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/qb_{sys.argv[1]}.json"))
print("QB cluster", d["cluster"], "ms/step", round(d["ms_per_step"], 1), "frac", round(d["roofline"]["frac"], 3), "TF", round(d["roofline"]["achieved"]), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
