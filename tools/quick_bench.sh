#!/bin/bash
# usage: tools/quick_bench.sh ROWS "cluster list" [workload]  -> one compact line per run
# a list entry may be cluster:rows to size the run to that cluster size's resident CTA count
ROWS=${1:-37888}; CL=${2:-"1 2"}; WL=${3:-cfg5}
mkdir -p gpurun_out
for e in $CL; do
  c=${e%%:*}; r=$ROWS; [[ "$e" == *:* ]] && r=${e##*:}
  timeout 300 python bench.py --workload $WL --rows $r --steps 2 --warmup 1 --no-cpu --no-e2e --cluster $c > gpurun_out/qb_$c.json 2> gpurun_out/qb_$c.err || tail -3 gpurun_out/qb_$c.err
  python - "$c" "$r" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/qb_{sys.argv[1]}.json"))
    print("QB cluster", d["cluster"], "rows", sys.argv[2], "ms/step", round(d["ms_per_step"], 1), "frac", round(d["roofline"]["frac"], 3), "TF", round(d["roofline"]["achieved"]), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("QB cluster", sys.argv[1], "failed", e)
PY
done
