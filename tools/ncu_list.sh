#!/bin/bash
# launch list (gpu__time_duration.sum per kernel, serialised, cold) of a command: tools/ncu_list.sh TAG cmd...
TAG=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv "$@" > gpurun_out/ncu_$TAG.log 2>&1
python - $TAG <<'PY'
import csv, sys
from collections import defaultdict
tag = sys.argv[1]
rows = list(csv.reader(l for l in open(f"gpurun_out/launches_{tag}.csv") if l.startswith('"')))
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r[ki].split("(")[0][:90]
    agg[k][0] += 1
    agg[k][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"LIST {tag}: total {tot/1e6:.3f} ms over {len(rows)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"LIST {v[1]/1e6:9.3f} ms {100*v[1]/tot:6.2f}%  x{v[0]:<4d} {k}")
PY
