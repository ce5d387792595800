"""Aggregate device-to-host bandwidth of N ranks copying at the same time into pinned host memory (the ceiling of the fp32 e2e
leg at N GPUs):   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/d2h_probe_multi.py [bind]
`bind` pins every rank to its GPU's NUMA node first (what bench.py does)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
bind = len(sys.argv) > 1 and sys.argv[1] == "bind"
numa = bench.bind_to_gpu_numa_node(local) if bind else None
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 29   # 2 GiB of fp32 per rank
dev = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
host = torch.empty(n, dtype=torch.float32, pin_memory=True)
host.zero_()
best = 1e9
for rep in range(4):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best = min(best, float(t.item()))
mine = torch.tensor([4.0 * n / dt / 1e9], dtype=torch.float64, device="cuda")
allv = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allv, mine)
if rank == 0:
    print(f"D2H x{world} bind={bind} numa={numa}: aggregate {world * 4.0 * n / best / 1e9:.1f} GB/s (slowest rank decides), "
          f"last rep per rank {[round(float(v.item()), 1) for v in allv]} GB/s", flush=True)
dist.destroy_process_group()
