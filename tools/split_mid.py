"""Column split with clusters of 2 / 4 for launches of more row tiles than clusters of 8 fit: automatic choice vs the pair flows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import random_modules
from sdrm_b200 import _lib
from sdrm_b200.engine import SamplerEngine
from sdrm_b200.models import make_schedule
for (L, H, I, T, nh, n) in [(200, 256, 3000, 60, 3, 1800), (128, 128, 2000, 50, 2, 1800), (96, 128, 1000, 40, 1, 1000), (830, 930, 1008, 83, 2, 2560), (830, 930, 1008, 83, 2, 4200), (830, 930, 1008, 83, 2, 9000), (950, 1000, 20000, 178, 4, 9000), (340, 490, 3125, 78, 1, 9000)]:
    diff, vae = random_modules(I, H, L, T, nh, seed=3, device="cuda")
    eng = SamplerEngine(); eng.pack_denoiser(diff, make_schedule(T, device="cuda"), 1.0); eng.pack_decoder(vae)
    out = torch.empty((n, I), dtype=torch.float32, device="cuda")
    res = []
    for no_split in (0, 1):
        eng.set_option(_lib.OPT_NO_SPLIT, no_split)
        ms = []
        for i in range(4):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.sample(n, seed=10 + i, out=out); b.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        res.append((min(ms), eng.lib.sdrm_last_split_size(eng.handle), eng.lib.sdrm_last_resident_mode(eng.handle)))
    print(f"MID L={L} I={I} T={T} nh={nh} n={n} ({(n + 127) // 128} tiles): split S={res[0][1]} {res[0][0]:.3f} ms | pair flows (resident {res[1][2]}) {res[1][0]:.3f} ms", flush=True)
