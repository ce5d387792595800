"""Dump the per-role event timeline of CTA 0 for a short cfg5-shaped run (debug tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, build_models
from sdrm_b200 import _lib
from sdrm_b200.train_SDRM import sample_ddpm, engine_for
w = dict(WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg5"]); w["T"] = 4
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 18944
diff, vae = build_models(w, "cuda")
eng = engine_for(diff, "cuda")   # needs a -DSDRM_TRACE (and, for the flags, -DSDRM_PERF_DEBUG) build selected with SDRM_B200_LIB
eng.set_option(_lib.OPT_CLUSTER, int(sys.argv[3]) if len(sys.argv) > 3 else 1)
eng.set_option(_lib.OPT_DEBUG_FLAGS, int(os.environ.get('SDRM_DEBUG_FLAGS', '0')))
eng.set_option(_lib.OPT_SUBTILES, int(os.environ.get('SDRM_SUBTILES', '0')))
eng.set_option(_lib.OPT_RESIDENT, int(os.environ.get('SDRM_NO_RESIDENT', '0')))
CAP = 8192
buf = torch.zeros(3 * CAP, dtype=torch.int64, device="cuda")
out = sample_ddpm(rows, diff, vae, w["L"], w["nd"], n_timesteps=w["T"], seed=1)
torch.cuda.synchronize()
eng.set_option(_lib.OPT_TRACE_BUFFER, buf.data_ptr())
out = sample_ddpm(rows, diff, vae, w["L"], w["nd"], n_timesteps=w["T"], seed=2)
torch.cuda.synchronize(); eng.set_option(_lib.OPT_TRACE_BUFFER, 0)
ev = buf.cpu().numpy().astype("uint64").reshape(3, CAP)
names = {0: {1: "P.layer_begin", 2: "P.act_ready", 3: "P.chunk_issued", 4: "P.empty_ok", 5: "P.issued", 6: "P.chunk_ok"}, 1: {1: "M.chunk_begin", 2: "M.acc_free", 3: "M.first_full", 4: "M.chunk_committed", 5: "M.full_ok", 6: "M.kb_issued", 7: "M.a_ready"},
         2: {1: "E.wait", 2: "E.acc_full", 3: "E.chunk_done", 4: "E.noise_done", 5: "E.layer_done"}}
allv = []
for role in range(3):
    for v in ev[role]:
        if v == 0: break
        allv.append((int(v) & ((1 << 40) - 1), role, int(v) >> 56, (int(v) >> 40) & 0xFFFF))
allv.sort()
t0 = allv[0][0]
limit = int(sys.argv[4]) if len(sys.argv) > 4 else 400
skip = int(sys.argv[5]) if len(sys.argv) > 5 else 0
for t, role, code, seq in allv[skip:skip + limit]:
    print(f"{(t - t0) / 1000.0:10.2f} us  {'   ' * role * 6}{names[role][code]} #{seq}")
