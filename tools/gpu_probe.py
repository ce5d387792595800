"""GPU bring-up probe for the tcgen05 layer engine: one dense layer vs torch, with diagnostics.

Run on a B200 box:  python tools/gpu_probe.py > gpurun_out/probe.log 2>&1
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdrm_b200 import _lib


def probe(M, K, N, split3, seed=0, structured=False, verbose=False):
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(seed)
    if structured:
        A = torch.zeros(M, K, device="cuda")
        A[torch.arange(M), torch.arange(M) % K] = 1.0
        W = (torch.arange(N, device="cuda").float()[:, None] * 1.0 + torch.arange(K, device="cuda").float()[None, :] / 256.0)
        bias = torch.zeros(N, device="cuda")
    else:
        A = torch.randn(M, K, device="cuda", generator=g)
        W = torch.randn(N, K, device="cuda", generator=g) / (K ** 0.5)
        bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda")
    ws_bytes = lib.sdrm_probe_linear_workspace_bytes(M, K, N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t0 = time.time()
    rc = lib.sdrm_probe_linear(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(out), M, K, N, int(split3),
                               _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
    _lib.check(rc, "sdrm_probe_linear")
    torch.cuda.synchronize()
    dt = time.time() - t0
    err_word = int(ws[:4].view(torch.int32).item())
    if split3:
        ref = (A.double() @ W.double().T + bias.double()).float()
    else:
        ref = (A.bfloat16().double() @ W.bfloat16().double().T + bias.double()).float()
    diff = (out - ref).abs()
    nan = int(torch.isnan(out).sum().item())
    mx = float(torch.nan_to_num(diff, nan=1e30).max().item())
    rel = float((torch.nan_to_num(out - ref).norm() / ref.norm()).item())
    ok = nan == 0 and mx < (2e-4 if split3 else 2e-3) * max(1.0, float(ref.abs().max()))
    print(f"probe M={M} K={K} N={N} split3={int(split3)} structured={int(structured)}: max_abs={mx:.3e} rel_fro={rel:.3e} "
          f"nan={nan} wd={err_word} t={dt*1e3:.1f}ms {'OK' if ok else 'MISMATCH'}", flush=True)
    if (not ok or verbose) and structured:
        torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
        print("out[0:10, 0:8]\n", out[:10, :8].cpu())
        print("ref[0:10, 0:8]\n", ref[:10, :8].cpu())
        print("out[64:72, 0:8]\n", out[64:72, :8].cpu())
    if not ok and not structured:
        bad = (diff > 1e-2).nonzero()
        print("  first bad idx:", bad[:10].tolist(), " count", bad.shape[0])
        rows_bad = torch.unique(bad[:, 0])[:20].tolist()
        cols_bad = torch.unique(bad[:, 1])[:20].tolist()
        print("  bad rows:", rows_bad, " bad cols:", cols_bad)
    return ok


def main():
    print(torch.cuda.get_device_name(0), torch.version.cuda, flush=True)
    results = []
    results.append(probe(128, 64, 64, False, structured=True, verbose=True))
    results.append(probe(128, 64, 64, False))
    results.append(probe(128, 64, 256, False))
    results.append(probe(128, 128, 64, False))
    results.append(probe(128, 48, 48, False))
    results.append(probe(100, 40, 40, False))
    results.append(probe(128, 256, 512, False))
    results.append(probe(300, 200, 300, False))
    results.append(probe(1000, 950, 950, False))
    results.append(probe(1000, 950, 950, True))
    results.append(probe(128 * 148 * 2 + 5, 340, 490, False))
    results.append(probe(2000, 1000, 20000, True))
    print("ALL OK" if all(results) else "SOME FAILED", flush=True)
    return 0 if all(results) else 1


if __name__ == "__main__":
    sys.exit(main())
