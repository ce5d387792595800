#!/bin/bash
# DRAM bytes of one engine wave at small grids: does the write-back traffic vanish when the scratch fits the L2?
M=dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second
for g in ${1:-"32 48 64 80"}; do
  ncu --metrics $M --clock-control none -k regex:layer_engine -c 1 --csv --log-file gpurun_out/r2_l2sweep_$g.csv python tools/l2_fit_probe.py $g 1 > /dev/null 2>&1
  python - $g <<'PY'
import csv, sys
g = int(sys.argv[1])
rows = [r for r in csv.reader(l for l in open(f"gpurun_out/r2_l2sweep_{g}.csv") if l.startswith('"'))]
d = {r[-3]: float(r[-1].replace(",", "")) for r in rows[1:]}
print(f"L2SWEEP grid {g}: dram read {d['dram__bytes_read.sum']/g/1e6:.0f} MB/CTA write {d['dram__bytes_write.sum']/g/1e6:.0f} MB/CTA  hit {d['lts__t_sector_hit_rate.pct']:.1f}%  "
      f"{d['gpu__time_duration.sum']/1e6:.2f} ms  clk {d['sm__cycles_elapsed.avg.per_second']/1e9:.3f} GHz  tensor {d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']:.1f}%")
PY
done
