"""Time sdrm_sample for an arbitrary denoiser shape under the per-handle options (A/B of sub-tiles / resident mode):
   python tools/shape_probe.py L H I T nh n"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from helpers import random_modules
from sdrm_b200 import _lib
from sdrm_b200.engine import SamplerEngine
from sdrm_b200.models import make_schedule

L, H, I, T, nh, n = (int(a) for a in sys.argv[1:7])
diff, vae = random_modules(I, H, L, T, nh, seed=3, device="cuda")
eng = SamplerEngine()
eng.pack_denoiser(diff, make_schedule(T, device="cuda"), 1.0)
eng.pack_decoder(vae)
out = torch.empty((n, I), dtype=torch.float32, device="cuda")
for sub, res in ((0, 0), (1, 0), (2, 0), (1, 1)):
    eng.set_option(_lib.OPT_SUBTILES, sub)
    eng.set_option(_lib.OPT_RESIDENT, res)
    ms = []
    for i in range(4):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.sample(n, seed=10 + i, out=out)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    print(f"SHAPE L={L} H={H} I={I} T={T} nh={nh} n={n}: subtiles {sub} resident_off {res} -> resident {eng.lib.sdrm_last_resident_mode(eng.handle)}"
          f"  ms {min(ms):.3f}  users/s {n / min(ms) * 1e3:.0f}", flush=True)
