"""K4 timing on the cfg 5 shard shape: each radix-select pass and the bit-pack pass read the [n, I] fp32 matrix once
(algorithmic bytes = 4 n I; the pack pass additionally writes n I / 8).  Prints GB/s against the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdrm_b200 import _lib
from sdrm_b200.sparsify import equal_sparsity_device, quantile_device

rows, cols = int(sys.argv[1]) if len(sys.argv) > 1 else 125000, int(sys.argv[2]) if len(sys.argv) > 2 else 20000
peak = 6535.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p)).get("hbm_gbs", peak)
lib = _lib.load()
x = torch.empty(rows, cols, device="cuda").normal_(-3.0, 2.5)
h = torch.zeros(2048, dtype=torch.int64, device="cuda")
bits = torch.empty(rows, (cols + 31) // 32, dtype=torch.int32, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
thr = float(quantile_device(x, 0.99))
key = 0
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
from sdrm_b200.sparsify import float_to_key
k = float_to_key(thr)
res = {}
for name, (prefix, pb, shift, nb) in {"pass1 (top 11 bits, skewed)": (0, 0, 21, 11), "pass2": (k >> 21, 11, 10, 11), "pass3": (k >> 10, 22, 0, 10)}.items():
    ms = timed(lambda: lib.sdrm_key_histogram(_lib.ptr(x), rows, cols, cols, prefix, pb, shift, nb, _lib.ptr(h), _lib.stream_ptr()))
    res[name] = (ms, 4.0 * rows * cols / ms / 1e6)
ms = timed(lambda: lib.sdrm_threshold_pack(_lib.ptr(x), rows, cols, cols, thr, 0, _lib.ptr(bits), bits.shape[1], _lib.ptr(cnt), _lib.stream_ptr()))
res["threshold+pack"] = (ms, (4.0 + 1 / 8) * rows * cols / ms / 1e6)
import time
torch.cuda.synchronize(); t0 = time.perf_counter(); pm = equal_sparsity_device(x, 0.99); torch.cuda.synchronize(); t1 = time.perf_counter()
for k_, (ms, gbs) in res.items():
    print(f"K4 {k_:32s} {ms:8.3f} ms  {gbs:8.1f} GB/s  {gbs / peak:5.2f} of measured HBM peak ({peak} GB/s)")
print(f"K4 equal_sparsity_device end to end ({rows}x{cols}, device walk, one state read): {(t1 - t0) * 1e3:.1f} ms, ones fraction {int(pm.ones) / x.numel():.6f}")
