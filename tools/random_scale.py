"""Multi-resolution sampling in the throughput regime (one wave of 148 row tiles at the cfg-5 shape): CTA pairs vs single CTAs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import WORKLOADS, build_models
from sdrm_b200 import _lib
from sdrm_b200.train_SDRM import sample_ddpm, engine_for
w = dict(WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg5"])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128
diff, vae = build_models(w, "cuda")
eng = engine_for(diff, "cuda")
out = torch.empty(n, w["I"], device="cuda")
ref = None
for cl in (2, 1, 0):
    eng.set_option(_lib.OPT_CLUSTER, cl)
    best = 1e9
    for rep in range(4):
        np.random.seed(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sample_ddpm(n, diff, vae, w["L"], w["nd"], timesteps="random", n_timesteps=w["T"], seed=5, out=out, reuse_packed=True)
        e1.record(); torch.cuda.synchronize()
        if rep: best = min(best, e0.elapsed_time(e1))
    same = True if ref is None else torch.equal(out, ref)
    if ref is None: ref = out.clone()
    print(f"RANDOM {sys.argv[1] if len(sys.argv) > 1 else 'cfg5'} rows {n}: cluster option {cl} -> cluster {_lib.load().sdrm_last_cluster_size(eng.handle)}: {best:.2f} ms, {n / best * 1e3:.0f} users/s, bit-identical {same}", flush=True)
