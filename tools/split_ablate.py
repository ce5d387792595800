"""Column-split mode at the cfg-1 shape under the load-ablation switches of a -DSDRM_PERF_DEBUG build (results wrong by design)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, build_models
from sdrm_b200 import _lib
from sdrm_b200.train_SDRM import sample_ddpm, engine_for
w = dict(WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg1"])
diff, vae = build_models(w, "cuda")
eng = engine_for(diff, "cuda")
for flags in [int(a) for a in sys.argv[2:]] or [0]:
    eng.set_option(_lib.OPT_DEBUG_FLAGS, flags)
    for _ in range(3):
        sample_ddpm(w["n"], diff, vae, w["L"], w["nd"], n_timesteps=w["T"], seed=1, reuse_packed=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        sample_ddpm(w["n"], diff, vae, w["L"], w["nd"], n_timesteps=w["T"], seed=1, reuse_packed=True)
    e1.record(); torch.cuda.synchronize()
    print(f"flags {flags}: {e0.elapsed_time(e1) / 10:.3f} ms per call, split {_lib.load().sdrm_last_split_size(eng.handle)}")
