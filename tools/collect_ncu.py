"""profiles/ncu_r02.txt + ncu_launches_r02.csv + ncu_traffic.json from the captures in gpurun_out/ (tools/collect_profiles.py
without the parts that re-assemble the bench lines and the K1 experiment log)."""
import csv, json, os, shutil, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
G, P, TAG = "gpurun_out", "profiles", "r02"
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "collect_profiles.py")).read()
ns = {}
exec(src[:src.index("txt = launch_list(")], ns)
launch_list, raw_metrics, WANT = ns["launch_list"], ns["raw_metrics"], ns["WANT"]
txt = launch_list(f"{G}/launches_{TAG}.csv", "ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-secondary` (cfg 5 shard)")
k1, m1 = raw_metrics(f"{G}/prof_{TAG}.ncu-rep", WANT, "ncu --set full, sdrm_layer_engine_kernel<2, false> (streaming flow), ONE full 148-CTA wave (18 944 users, cfg 5), discard warp on")
txt += k1
for rep, title in ((f"{G}/prof_{TAG}_res.ncu-rep", "ncu --set full, sdrm_layer_engine_kernel<2, true> (resident flow), cfg 2: 5 429 users x 3 125 items, T = 78, one row tile per CTA"),
                   (f"{G}/prof_k6b.ncu-rep", "ncu --set full, sdrm_small_chain_kernel<5> (K6), cfg 3: 9 558 users x 8 582 items, T = 93 (first r02 session; kernel unchanged)"),
                   (f"{G}/prof_{TAG}_gemm.ncu-rep", "ncu --set full, sdrm_gemm_pair_kernel (K5), one GEMM of the training step at the cfg-5 layer shape (49 152 rows) (first r02 session; kernel unchanged)"),
                   (f"{G}/prof_{TAG}_topk.ncu-rep", "ncu --set full, topk_pool_f32_kernel<true> (K3), k = 50, 65 536 x 20 000 fp32 scores")):
    if os.path.exists(rep):
        txt += raw_metrics(rep, WANT, title)[0]
open(f"{P}/ncu_{TAG}.txt", "w").write("\n".join(txt) + "\n")
shutil.copyfile(f"{G}/launches_{TAG}.csv", f"{P}/ncu_launches_{TAG}.csv")
users = 18944
raw = list(csv.reader(subprocess.run(["ncu", "-i", f"{G}/prof_{TAG}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
unit = dict(zip(raw[0], raw[1]))
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
rd = float(m1["dram__bytes_read.sum"]) * scale[unit["dram__bytes_read.sum"]]
wr = float(m1["dram__bytes_write.sum"]) * scale[unit["dram__bytes_write.sum"]]
json.dump({"dram_bytes_per_user": (rd + wr) / users, "algorithmic_hbm_bytes_per_user": 80000,
           "source": f"ncu --set full, sdrm_layer_engine_kernel<2>, {users} users (one 148-CTA wave), cfg5, {TAG}: dram__bytes_read.sum {rd / 1e9:.1f} GB + dram__bytes_write.sum {wr / 1e9:.1f} GB; scaled per user",
           "tensor_pipe_active_pct": float(m1["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]),
           "sm_ghz": float(m1["sm__cycles_elapsed.avg.per_second"]), "kernel_ms": float(m1["gpu__time_duration.sum"])}, open(f"{P}/ncu_traffic.json", "w"), indent=1)
print(open(f"{P}/ncu_{TAG}.txt").read())
