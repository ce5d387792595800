#!/bin/bash
# A/B of tuning builds (sdrm_b200/csrc/build.py --out=libX.so -D...): tools/ab_variants.sh ROWS "lib:subtiles ..."
ROWS=${1:-125000}; LIST=${2:-"libsdrm_b200.so:1"}
mkdir -p gpurun_out
for e in $LIST; do
  lib=${e%%:*}; s=${e##*:}
  SDRM_B200_LIB=$PWD/sdrm_b200/csrc/$lib timeout 300 python bench.py --rows $ROWS --steps 3 --warmup 2 --no-cpu --no-e2e --subtiles $s > gpurun_out/abv.json 2> gpurun_out/abv.err || tail -3 gpurun_out/abv.err
  python - "$lib" "$s" "$ROWS" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/abv.json"))
    print("ABV", sys.argv[1], "subtiles", sys.argv[2], "rows", sys.argv[3], "ms/step", round(d["ms_per_step"], 1), "users/s", round(d["value"]), "frac", round(d["roofline"]["frac"], 3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("ABV", sys.argv[1], "failed", e)
PY
done
