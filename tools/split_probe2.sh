#!/bin/bash
# the narrower dataset configurations: automatic choice (column split where it fits) against the pair flows (--cluster 2: resident / streaming)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sampler_gpu.py -x -q -k "column_split or resident_mode" 2>&1 | tail -5
for wl in cfg2 cfg4 cfg1; do for c in 0 2; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --no-secondary --no-e2e --cluster $c 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl cluster $c', round(d['ms_per_step'],3), 'ms', round(d['value']), 'users/s')"
done; done
