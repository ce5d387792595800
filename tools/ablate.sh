#!/bin/bash
# energy / critical-path ablations of the engine kernel under the board power cap (results of flags != 0 are numerically wrong)
#   tools/ablate.sh ROWS   (needs sdrm_b200/csrc/libdbg.so = build.py --out=libdbg.so -DSDRM_PERF_DEBUG)
ROWS=${1:-56832}
mkdir -p gpurun_out
run() {  # label, lib, flags, extra bench args
  SDRM_DEBUG_FLAGS=$3 SDRM_B200_LIB=$PWD/sdrm_b200/csrc/$2 timeout 300 python bench.py --rows $ROWS --steps 3 --warmup 2 --no-cpu --no-e2e $4 > gpurun_out/abl.json 2> gpurun_out/abl.err || tail -3 gpurun_out/abl.err
  python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/abl.json"))
    print("ABL", sys.argv[1], "ms/step", round(d["ms_per_step"], 1), "users/s", round(d["value"]), "frac", round(d["roofline"]["frac"], 3), "cluster", d["cluster"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("ABL", sys.argv[1], "failed", e)
PY
}
run base libdbg.so 0 ""
run noise-no-rng libdbg.so 16 ""
run noise-no-state libdbg.so 32 ""
run act-store-fixed libdbg.so 64 ""
run no-store-wait libdbg.so 128 ""
run no-act-stores libdbg.so 1 ""
run base2 libdbg.so 0 ""
