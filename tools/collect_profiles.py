"""Assemble profiles/*_r02.* from gpurun_out/ (raw logs of this round).  Run after the GPU calls; everything it writes is committed."""
import csv, glob, json, os, shutil, subprocess, sys
from collections import defaultdict
TAG = "r02"
G, P = "gpurun_out", "profiles"


def cp(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))


def launch_list(path, title):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = r[ki].split("(")[0][:100]
        agg[k][0] += 1
        agg[k][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    out = [f"# {title}: gpu__time_duration.sum per kernel, --clock-control none (serialised, cold caches: shares, not absolutes)",
           f"total {tot / 1e6:.3f} ms over {len(rows)} launches", ""]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
        out.append(f"{v[1] / 1e6:10.3f} ms {100 * v[1] / tot:6.2f}%  x{v[0]:<4d} {k}")
    return out


def raw_metrics(rep, want, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    names, units, vals = r[0], r[1], r[2]
    out = ["", f"# {title}"]
    for n, u, v in zip(names, units, vals):
        if n in want:
            out.append(f"{n:82s} {v} {u}")
    return out, dict(zip(names, vals))


WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active",
        "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio"]

txt = launch_list(f"{G}/launches_{TAG}.csv", "ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e` (cfg 5 shard)")
k1, m1 = raw_metrics(f"{G}/prof_{TAG}.ncu-rep", WANT, "ncu --set full, sdrm_layer_engine_kernel<2>, ONE full 148-CTA wave (18 944 users, cfg 5), discard warp on")
txt += k1
if os.path.exists(f"{G}/prof_k6b.ncu-rep"):
    txt += raw_metrics(f"{G}/prof_k6b.ncu-rep", WANT, "ncu --set full, sdrm_small_chain_kernel<5> (K6), cfg 3: 9 558 users x 8 582 items, T = 93")[0]
if os.path.exists(f"{G}/prof_{TAG}_gemm.ncu-rep"):
    txt += raw_metrics(f"{G}/prof_{TAG}_gemm.ncu-rep", WANT, "ncu --set full, sdrm_gemm_pair_kernel (K5), one GEMM of the training step at the cfg-5 layer shape (49 152 rows)")[0]
open(f"{P}/ncu_{TAG}.txt", "w").write("\n".join(txt) + "\n")
shutil.copyfile(f"{G}/launches_{TAG}.csv", f"{P}/ncu_launches_{TAG}.csv")
open(f"{P}/ncu_launches_train_{TAG}.txt", "w").write("\n".join(launch_list(f"{G}/launches_train_{TAG}.csv",
     "ncu launch list of `python bench.py --mode train --steps 1 --warmup 0 --e2e-steps 1` (7 steps on our GEMMs: bf16x3, bf16, 5 e2e; 1 step on torch / cuBLAS fp32)")) + "\n")

# traffic per user for bench.py's roofline.traffic
users = 18944
rd, wr = float(m1["dram__bytes_read.sum"]), float(m1["dram__bytes_write.sum"])
unit = dict(zip(*[r for r in list(csv.reader(subprocess.run(["ncu", "-i", f"{G}/prof_{TAG}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))[:2]]))
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
rd *= scale[unit["dram__bytes_read.sum"]]; wr *= scale[unit["dram__bytes_write.sum"]]
json.dump({"dram_bytes_per_user": (rd + wr) / users, "algorithmic_hbm_bytes_per_user": 80000,
           "source": f"ncu --set full, sdrm_layer_engine_kernel<2>, {users} users (one 148-CTA wave), cfg5, {TAG}: dram__bytes_read.sum {rd / 1e9:.1f} GB + dram__bytes_write.sum {wr / 1e9:.1f} GB; scaled per user",
           "tensor_pipe_active_pct": float(m1["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]),
           "sm_ghz": float(m1["sm__cycles_elapsed.avg.per_second"]), "kernel_ms": float(m1["gpu__time_duration.sum"])}, open(f"{P}/ncu_traffic.json", "w"), indent=1)

for c in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
    cp(f"r02a_bench_{c}_n1.json", f"bench_{c}_n1_{TAG}.json")
cp("r02a_bench_cfg5_reference.json", f"bench_cfg5_reference_{TAG}.json")
cp("r02a_bench_train_n1.json", f"bench_train_n1_{TAG}.json")
cp("r2_train_n2.json", f"bench_train_n2_{TAG}.json")
cp("r02a_k2k3.txt", f"k2_k3_{TAG}.txt")
cp("r02a_k4.txt", f"k4_sparsify_{TAG}.txt")
cp("r02a_gpu_tests.log", f"gpu_tests_{TAG}.log")
cp("r2_tests_2gpu.log", f"gpu_tests_2gpu_{TAG}.log")
with open(f"{P}/sass_{TAG}.txt", "w") as fh:
    fh.write(subprocess.run([sys.executable, "tools/sass_summary.py"], capture_output=True, text=True).stdout)

# K1 bound analysis: the raw lines of the experiments quoted in DESIGN.md section 4
parts = [("# tools/l2_fit_probe.py GRID 2: one cfg-5 wave on GRID CTAs (CUDA events)", "r2_l2fit.log"),
         ("# tools/l2_sweep.sh: ncu DRAM bytes of one wave at small grids", "r2_l2sweep.log"),
         ("# tools/ablate.sh 56832 (first set: flags 4 / 1 / 5, cluster 4 / 8 / 2)", "r2_ablate.log"),
         ("# tools/ablate.sh 56832 (second set: flags 16 / 32 / 64 / 128 / 1)", "r2_ablate2.log"),
         ("# tools/ab_variants.sh 125000: one (libv_s1) vs two (libsdrm_b200) store slots, 6 vs 5 (n5) stages", "r2_ab1.log"),
         ("# tools/ab_discard.sh 125000: SDRM_NO_DISCARD=1 vs 0", "r2_abd.log"),
         ("# ncu, one wave, discard off / on (tools/ncu_wave.sh)", None),
         ("# tools/ubench_discard MODE NSLAB 400 under ncu (mode 0 none, 1 discard after the last read, 2 rotating slabs)", "r2_discard.log")]
out = []
for title, f in parts:
    out.append(title)
    if f and os.path.exists(f"{G}/{f}"):
        out += [l.rstrip() for l in open(f"{G}/{f}") if l.strip()]
    if f is None:
        out += ["WAVE nodiscard: 34.17 ms  clk 1.263 GHz  tensor 75.9%  dram R 38.2 GB W 60.1 GB  L2 hit 84.5%  L2 rd 350 GB wr 69 GB  inst 10.59 G",
                "WAVE discard: 34.26 ms  clk 1.313 GHz  tensor 72.8%  dram R 33.3 GB W 41.7 GB  L2 hit 84.9%  L2 rd 350 GB wr 69 GB  inst 10.69 G"]
    out.append("")
for g in (148, 112, 96):
    f = f"{G}/r2_l2fit_ncu_{g}.csv"
    if os.path.exists(f):
        rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
        d = {r[-3]: r[-1] for r in rows[1:]}
        out.append(f"ncu grid {g}: " + "  ".join(f"{k.split('.')[0]}={v}" for k, v in d.items()))
open(f"{P}/k1_bound_{TAG}.txt", "w").write("\n".join(out) + "\n")
print("profiles written:", sorted(f for f in os.listdir(P) if TAG in f))
