"""D2H bandwidth into pinned host memory with 1, 2 and 4 concurrent copy streams (what bounds the fp32 e2e path)."""
import time, torch
n = 1 << 30   # 4 GiB of fp32
dev = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
host = torch.empty(n, dtype=torch.float32, pin_memory=True)
for ns in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    part = n // ns
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                host[i * part:(i + 1) * part].copy_(dev[i * part:(i + 1) * part], non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"D2H {ns} stream(s): {4 * n / best / 1e9:.1f} GB/s")
