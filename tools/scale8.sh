#!/bin/bash
# one multi-GPU box: host D2H ceiling, the sampler bench, cfg 4 sharded, the training step.  usage: tools/scale8.sh N
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tools/d2h_probe_multi.py bind 2>&1 | grep "D2H x"
$TR --master-port 29512 tools/d2h_probe_multi.py 2>&1 | grep "D2H x"
$TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
$TR --master-port 29514 bench.py --gpus $N --workload cfg4 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_cfg4_$N.json 2> gpurun_out/scale_cfg4_$N.err
$TR --master-port 29515 bench.py --gpus $N --mode train --steps 10 --warmup 3 > gpurun_out/scale_train_$N.json 2> gpurun_out/scale_train_$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for tag in ("scale", "scale_cfg4", "scale_train"):
    try:
        d = json.load(open(f"gpurun_out/{tag}_{n}.json"))
        print(tag.upper(), "n_gpus", d["n_gpus"], d["unit"], round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]),
              "frac", round(d["roofline"]["frac"], 3), {k: d["e2e"].get(k) for k in ("packed", "host_numa_binding") if k in d["e2e"]}, d["clocks"])
    except Exception as e:
        print(tag.upper(), n, "failed", e)
PY
