"""Does the engine's DRAM traffic come from L2 capacity?  One full wave of cfg-5 row tiles on GRID CTAs (SDRM_OPT_GRID_LIMIT):
   python tools/l2_fit_probe.py GRID [waves]            -> ms per wave (CUDA events)
   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,\
gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:layer_engine python tools/l2_fit_probe.py GRID 1
Hot scratch per CTA at cfg 5 = 0.98 MB, so GRID = 148 -> 145 MB, 112 -> 110 MB, 96 -> 94 MB (L2 = 126 MB)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from sdrm_b200 import _lib
from sdrm_b200.train_SDRM import engine_for, sample_ddpm

grid = int(sys.argv[1])
waves = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = bench.WORKLOADS["cfg5"]
dev = torch.device("cuda", 0)
diff, vae = bench.build_models(w, dev)
engine_for(diff, dev).set_option(_lib.OPT_GRID_LIMIT, grid)
n = grid * 128 * waves
out = torch.empty((n, w["I"]), dtype=torch.float32, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for i in range(3):
    ev[i].record()
    sample_ddpm(n, diff, vae, w["L"], w["nd"], n_timesteps=w["T"], seed=10 + i, out=out)
ev[3].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
print(f"L2FIT grid {grid} waves {waves} rows {n}: ms per call {[round(m, 2) for m in ms]}  ms per wave {min(ms) / waves:.2f}  "
      f"users/s per CTA {n / (min(ms) * 1e-3) / grid:.0f}")
