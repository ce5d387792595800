#!/bin/bash
# same-box A/B of library builds / environment switches: tools/ab_libs.sh ROWS "label:lib:ENV=1 ..."
ROWS=${1:-125000}; shift
for e in "$@"; do
  IFS=: read label lib envs <<< "$e"
  env $envs SDRM_B200_LIB=$PWD/sdrm_b200/csrc/$lib timeout 300 python bench.py --rows $ROWS --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/abl2.json 2> gpurun_out/abl2.err || tail -3 gpurun_out/abl2.err
  python - "$label" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/abl2.json"))
    print("ABL2", sys.argv[1], "ms/step", round(d["ms_per_step"], 1), "users/s", round(d["value"]), "frac", round(d["roofline"]["frac"], 3), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("ABL2", sys.argv[1], "failed", e)
PY
done
