#!/bin/bash
# Everything a round's numbers come from, on ONE GPU box: tools/round_validation.sh [TAG]   -> gpurun_out/<TAG>_*
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/${TAG}_gpu_tests.log; cat $O/${TAG}_gpu_tests.log | tail -2
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_cfg5_reference.json 2> $O/${TAG}_bench_ref.err
python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_cfg5_n1.json 2> $O/${TAG}_bench_cfg5.err || tail -3 $O/${TAG}_bench_cfg5.err
for c in cfg1 cfg2 cfg3 cfg4; do
  python bench.py --workload $c --steps 20 --warmup 3 > $O/${TAG}_bench_${c}_n1.json 2> $O/${TAG}_bench_$c.err || tail -3 $O/${TAG}_bench_$c.err
done
python bench.py --mode train --steps 10 --warmup 3 > $O/${TAG}_bench_train_n1.json 2> $O/${TAG}_bench_train.err || tail -3 $O/${TAG}_bench_train.err
python tools/bench_train_kernels.py > $O/${TAG}_k2k3.txt 2>&1
python tools/bench_sparsify.py 125000 20000 > $O/${TAG}_k4.txt 2>&1
python - $TAG <<'PY'
import json, sys, glob
tag = sys.argv[1]
for f in sorted(glob.glob(f"gpurun_out/{tag}_bench_*.json")):
    try:
        d = json.load(open(f))
        r = d.get("roofline") or {}
        print("VAL", f.split("/")[-1], "value", round(d["value"]), d["unit"], "ms", round(d.get("ms_per_step", 0), 3), "frac", round(r.get("frac", 0), 4) if r else None,
              "e2e", round((d.get("e2e") or {}).get("value", 0)), "clk", (d.get("clocks") or {}).get("sm_mhz"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("eager_cuda") or {}).get("value"))
    except Exception as e:
        print("VAL", f, "unreadable", e)
PY
