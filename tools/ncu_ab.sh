#!/bin/bash
# ncu counters of the engine kernel for both sub-tile settings (2 tiles per CTA): tools/ncu_ab.sh [ROWS]
ROWS=${1:-37888}
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum
for s in 1 2; do
  ncu --metrics $M --clock-control none -k regex:sdrm_layer_engine -c 1 --csv --log-file gpurun_out/ncu_ab_$s.csv \
    python bench.py --rows $ROWS --steps 1 --warmup 0 --no-cpu --no-e2e --subtiles $s > gpurun_out/ncu_ab_$s.log 2>&1
  python - $s <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(f"gpurun_out/ncu_ab_{sys.argv[1]}.csv")) if len(r) > 10]
h = rows[0]; i_n, i_v, i_u = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
print("subtiles", sys.argv[1], "; ".join(f"{r[i_n].split('.')[0]}={r[i_v]}{r[i_u]}" for r in rows[1:]))
PY
done
