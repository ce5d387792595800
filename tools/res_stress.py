"""Resident flow without cluster-scope release / fence: full-length launches compared bit for bit with the streaming flow."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import random_modules
from sdrm_b200 import _lib
from sdrm_b200.engine import SamplerEngine
from sdrm_b200.models import make_schedule
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
bad = 0
for (L, H, I, T, nh, n) in [(340, 490, 3125, 78, 1, 12000), (400, 550, 729, 43, 0, 19000), (200, 256, 1000, 60, 3, 12000), (512, 300, 500, 30, 1, 40000)]:
    diff, vae = random_modules(I, H, L, T, nh, seed=3, device="cuda")
    eng = SamplerEngine(); eng.pack_denoiser(diff, make_schedule(T, device="cuda"), 1.0); eng.pack_decoder(vae)
    for r in range(reps):
        eng.set_option(_lib.OPT_RESIDENT, 1)
        ref = eng.sample(n, seed=50 + r, check=True).clone()
        eng.set_option(_lib.OPT_RESIDENT, 0)
        out = eng.sample(n, seed=50 + r, check=True)
        res = eng.lib.sdrm_last_resident_mode(eng.handle)
        if not torch.equal(out, ref):
            bad += 1
            print("MISMATCH", L, n, r)
    print(f"L={L} T={T} n={n}: {reps} seeds, resident {res}: {'bit-identical' if not bad else 'MISMATCH'}", flush=True)
print("RES STRESS", "FAILED" if bad else "OK")
