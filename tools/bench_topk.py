"""K3 microbenchmark: python tools/bench_topk.py  (SDRM_TOPK_POOL_MIN_K=64 selects the insertion kernel for every k: A/B)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdrm_b200 import metrics

peak = 6535.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


tag = os.environ.get("SDRM_TOPK_POOL_MIN_K", "16")
for rows, items in ((65536, 20000), (9558, 8582), (5429, 3125)):
    sc = torch.randn(rows, items, device="cuda")
    for k in (10, 20, 50, 64):
        ms = timed(lambda: metrics.topk_device(sc, k))
        print(f"K3 topk k={k:<2d} {rows}x{items} fp32 (pool above k={tag})".ljust(60),
              f"{ms:8.3f} ms  {4.0 * rows * items / ms / 1e6:8.1f} GB/s  {4.0 * rows * items / ms / 1e6 / peak:5.2f} of measured HBM peak ({peak} GB/s)")
    del sc
