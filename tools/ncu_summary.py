"""Summarise gpurun_out/launches_rNN.csv and prof_rNN.ncu-rep into profiles/ (text, committed)."""
import csv, subprocess, sys
from collections import defaultdict
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = list(csv.reader(l for l in open(f"gpurun_out/launches_{tag}.csv") if l.startswith('"')))
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r[ki].split("(")[0][:100]
    agg[k][0] += 1
    agg[k][1] += float(r[vi])
tot = sum(v[1] for v in agg.values())
out = [f"# ncu launch list ({tag}): gpu__time_duration.sum per kernel, --clock-control none", f"total {tot/1e6:.3f} ms over {len(rows)} launches", ""]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:15]:
    out.append(f"{v[1]/1e6:10.3f} ms {100*v[1]/tot:6.2f}%  x{v[0]:<4d} {k}")
raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
names, units, vals = r[0], r[1], r[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active",
        "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__cluster_size"]
out += ["", f"# ncu --set full, kernel sdrm_layer_engine_kernel ({tag}), one launch"]
for n, u, v in zip(names, units, vals):
    if n in want:
        out.append(f"{n:75s} {v} {u}")
open(f"profiles/ncu_{tag}.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
