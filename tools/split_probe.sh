#!/bin/bash
# column-split mode: bit-identity test, then cfg 1 with the split on (automatic) and off (--cluster 2 = the previous pair flow)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sampler_gpu.py -x -q -k "column_split" 2>&1 | tail -15
for c in 0 2; do
  timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu --no-secondary --cluster $c > gpurun_out/split_cfg1_c$c.json 2> gpurun_out/split_cfg1_c$c.err || tail -3 gpurun_out/split_cfg1_c$c.err
  python - $c <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/split_cfg1_c{sys.argv[1]}.json"))
    print("cfg1 cluster", sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "users/s", round(d["value"]), "e2e", d.get("e2e", {}).get("value"))
except Exception as e:
    print("failed", e)
PY
done
