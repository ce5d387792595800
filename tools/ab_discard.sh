#!/bin/bash
# same-box A/B of the dead-buffer discard: tools/ab_discard.sh ROWS
ROWS=${1:-125000}
mkdir -p gpurun_out
for nd in 1 0 1 0; do
  SDRM_NO_DISCARD=$nd timeout 300 python bench.py --rows $ROWS --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/abd.json 2> gpurun_out/abd.err || tail -3 gpurun_out/abd.err
  python - "$nd" "$ROWS" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/abd.json"))
    print("ABD no_discard", sys.argv[1], "rows", sys.argv[2], "ms/step", round(d["ms_per_step"], 1), "users/s", round(d["value"]), "frac", round(d["roofline"]["frac"], 3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("ABD", sys.argv[1], "failed", e)
PY
done
