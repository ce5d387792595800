#!/bin/bash
# one full 148-CTA wave of cfg 5 under ncu (light metric set): tools/ncu_wave.sh TAG   (env is passed through, e.g. SDRM_NO_DISCARD=1)
M=dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,smsp__inst_executed.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum
TAG=${1:-x}
ncu --metrics $M --clock-control none -k regex:layer_engine -c 1 --csv --log-file gpurun_out/r2_wave_$TAG.csv python bench.py --rows 18944 --steps 1 --warmup 0 --no-cpu --no-e2e > /dev/null 2>&1
python - $TAG <<'PY'
import csv, sys
tag = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(f"gpurun_out/r2_wave_{tag}.csv") if l.startswith('"'))]
d = {r[-3]: float(r[-1].replace(",", "")) for r in rows[1:]}
print(f"WAVE {tag}: {d['gpu__time_duration.sum']/1e6:.2f} ms  clk {d['sm__cycles_elapsed.avg.per_second']/1e9:.3f} GHz  tensor {d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']:.1f}%  "
      f"dram R {d['dram__bytes_read.sum']/1e9:.1f} GB W {d['dram__bytes_write.sum']/1e9:.1f} GB  L2 hit {d['lts__t_sector_hit_rate.pct']:.1f}%  L2 rd {d['lts__t_sectors_srcunit_tex_op_read.sum']*32/1e9:.0f} GB wr {d['lts__t_sectors_srcunit_tex_op_write.sum']*32/1e9:.0f} GB  inst {d['smsp__inst_executed.sum']/1e9:.2f} G")
PY
