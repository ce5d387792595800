"""Engine-only GEMM throughput sweep through sdrm_probe_linear (kernel time = slope over repeat counts)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdrm_b200 import _lib

lib = _lib.load()

def run(M, K, N, split3=0, reps=(1, 9)):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.zeros(N, device="cuda"); out = torch.empty(M, N, device="cuda")
    wsb = lib.sdrm_probe_linear_workspace_bytes(M, K, N); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    ts = []
    for rep in reps:
        for _ in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _r in range(rep):
                _lib.check(lib.sdrm_probe_linear(_lib.ptr(A), _lib.ptr(W), _lib.ptr(b), _lib.ptr(out), M, K, N, split3, _lib.ptr(ws), wsb, _lib.stream_ptr()))
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ts.append(dt)
    t = (ts[1] - ts[0]) / (reps[1] - reps[0])
    passes = 3 if split3 else 1
    nch = (N + 255) // 256; nc = -(-(-(-N // nch)) // 16) * 16
    kb = -(-K // 64)
    tiles = -(-M // 128)
    per_cta_tiles = -(-tiles // 148)
    bytes_per_cta = per_cta_tiles * nch * passes * kb * (16384 + nc * 128)
    flops = 2.0 * M * K * N
    print(f"M={M} K={K} N={N} s3={split3}: {t*1e3:8.3f} ms  {flops/t/1e12:7.1f} TFLOP/s(alg)  NC={nc} L2->SM {bytes_per_cta/t/1e9:6.1f} GB/s/SM "
          f"= {bytes_per_cta/(t*1.9e9):5.1f} B/cyc@1.9GHz", flush=True)

M = 148 * 128 * 2
for K, N in [(960, 960), (960, 256), (960, 128), (960, 64), (960, 512), (1920, 960), (64, 256), (320, 256), (1024, 20000)]:
    run(M if N < 20000 else 148 * 128, K, N)
run(148 * 128, 1024, 20000, 1)
