"""Stress the column-split flow's cross-CTA hand-off (plain mbarrier arrivals after completed TMA stores): many launches at full
chain length, every one compared bit for bit with the one-CTA-per-tile flow of the same seed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, build_models
from sdrm_b200 import _lib
from sdrm_b200.engine import SamplerEngine
from sdrm_b200.models import make_schedule
lib = _lib.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
shapes = [dict(WORKLOADS["cfg1"]), dict(WORKLOADS["cfg4"]), dict(WORKLOADS["cfg2"]), dict(WORKLOADS["cfg5"], n=400, I=2000)]
bad = 0
for w in shapes:
    diff, vae = build_models(w, "cuda")
    eng = SamplerEngine()
    eng.pack_denoiser(diff, make_schedule(w["T"], device="cuda"), w["nd"])
    eng.pack_decoder(vae)
    n = w["n"]
    for r in range(reps):
        eng.set_option(_lib.OPT_NO_SPLIT, 1)
        ref = eng.sample(n, seed=100 + r, check=True).clone()
        eng.set_option(_lib.OPT_NO_SPLIT, 0)
        for cl in (0, 4):
            eng.set_option(_lib.OPT_CLUSTER, cl)
            out = eng.sample(n, seed=100 + r, check=True)
            s = lib.sdrm_last_split_size(eng.handle)
            if s and not torch.equal(out, ref):
                bad += 1
                print("MISMATCH", w["desc"], "rep", r, "cluster option", cl, "split", s, float((out - ref).abs().max()))
        eng.set_option(_lib.OPT_CLUSTER, 0)
    print(f"{w['desc']}: {reps} seeds x (automatic, clusters of 4) vs one CTA per tile: {'all bit-identical' if not bad else str(bad) + ' mismatches'}; last split size {s}", flush=True)
print("STRESS", "FAILED" if bad else "OK")
