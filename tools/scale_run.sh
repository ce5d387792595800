#!/bin/bash
# N = 1, 2, 4, 8 back to back (the driver's scaling run), one JSON line per N into gpurun_out/scale_<N>.json
mkdir -p gpurun_out
MAXN=${1:-8}
for N in 1 2 4 8; do
  [ $N -gt $MAXN ] && break
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  fi
  python - $N <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/scale_{sys.argv[1]}.json"))
    print("SCALE n_gpus", d["n_gpus"], "users/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), d["clocks"])
except Exception as e:
    print("SCALE", sys.argv[1], "failed", e)
PY
done
