"""Multi-resolution sampling (timesteps='random') at a dataset-sized configuration: column split on (automatic) / off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import WORKLOADS, build_models
from sdrm_b200 import _lib
from sdrm_b200.train_SDRM import sample_ddpm, engine_for
for name in sys.argv[1:] or ["cfg1"]:
    w = dict(WORKLOADS[name])
    diff, vae = build_models(w, "cuda")
    eng = engine_for(diff, "cuda")
    for no_split in (0, 1):
        eng.set_option(_lib.OPT_NO_SPLIT, no_split)
        best = 1e9
        for rep in range(6):
            np.random.seed(0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sample_ddpm(w["n"], diff, vae, w["L"], w["nd"], timesteps="random", n_timesteps=w["T"], seed=1, reuse_packed=True)
            e1.record(); torch.cuda.synchronize()
            if rep: best = min(best, e0.elapsed_time(e1))
        print(f"{name} random mode, no_split={no_split}: {best:.3f} ms per call ({w['n'] / best * 1e3:.0f} users/s), split {_lib.load().sdrm_last_split_size(eng.handle)}")
