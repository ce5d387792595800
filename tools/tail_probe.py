"""cfg 5 shard (125 000 users = 977 row tiles on 148 CTAs = 6.6 rounds): one launch vs full rounds in pair mode + the tail tiles on
column-split clusters (separate launches on the same stream; rows are keyed by their global id, so the bits do not change)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, build_models
from sdrm_b200 import _lib
from sdrm_b200.engine import SamplerEngine
from sdrm_b200.models import make_schedule
w = WORKLOADS["cfg5"]
diff, vae = build_models(w, "cuda")
eng = SamplerEngine(); eng.pack_denoiser(diff, make_schedule(w["T"], device="cuda"), w["nd"]); eng.pack_decoder(vae)
n = w["n"]
out = torch.empty((n, w["I"]), dtype=torch.float32, device="cuda")
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
def parts(cuts):
    def run():
        lo = 0
        for hi in cuts + [n]:
            if hi > lo:
                eng.sample(hi - lo, row_offset=lo, seed=7, out=out[lo:hi])
            lo = hi
    return run
one = timed(parts([]))
print(f"TAIL one launch: {one:.2f} ms", flush=True)
ref = out.clone()
full = 6 * 148 * 128
for cuts in ([full], [full, full + 74 * 128], [5 * 148 * 128, 5 * 148 * 128 + 148 * 128], [full, full + 45 * 128]):
    ms = timed(parts(list(cuts)))
    same = torch.equal(out, ref)
    print(f"TAIL cuts at rows {cuts}: {ms:.2f} ms ({one / ms:.3f} x), bit-identical {same}", flush=True)
