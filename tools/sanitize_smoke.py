"""Small launches of every kernel family and data flow in one process (a quick smoke after kernel edits; also the script to
put under `compute-sanitizer --tool memcheck` where the pool allows it)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from helpers import random_modules
from sdrm_b200 import _lib, metrics
from sdrm_b200.engine import SamplerEngine
from sdrm_b200.models import make_schedule
from sdrm_b200.sparsify import equal_sparsity_device

for (n, I, H, L, T, nh, opts) in ((300, 201, 96, 80, 3, 1, {}), (700, 729, 200, 264, 3, 2, {}),                  # resident, odd item counts
                                  (900, 403, 128, 120, 3, 1, {_lib.OPT_GRID_LIMIT: 4}),                          # two interleaved sub-tiles
                                  (520, 1008, 300, 600, 2, 1, {}), (130, 333, 40, 40, 3, 2, {})):               # streaming; K6
    diff, vae = random_modules(I, H, L, T, nh, seed=1, device="cuda")
    eng = SamplerEngine()
    eng.pack_denoiser(diff, make_schedule(T, device="cuda"), 1.0)
    eng.pack_decoder(vae)
    for k, v in opts.items():
        eng.set_option(k, v)
    out = eng.sample(n, seed=3, check=True)
    print("sample", n, I, L, "resident", eng.lib.sdrm_last_resident_mode(eng.handle), float(out.abs().mean()), flush=True)
    for k in (10, 50):
        if k <= I:
            metrics.topk_device(out, k)
    equal_sparsity_device(out, 0.9)
x = torch.randn(37, 1777, device="cuda", dtype=torch.float64)
metrics.topk_device(x, 10)
metrics.topk_device(x, 40)
torch.cuda.synchronize()
print("done")
