"""Does a CS = 4 launch (two tcgen05 pairs sharing one multicast weight stream; 132 of 148 SMs can hold such clusters) plus a
concurrent CS = 2 launch on the 16 SMs it leaves free beat the plain CS = 2 launch on all 148 SMs at cfg 5?
   python tools/hybrid_probe.py [rounds]
Prints ms and row tiles per ms for: CS = 2 alone (148 x rounds tiles), CS = 4 alone (132 x rounds tiles), and both kernels
together on two streams (132 x rounds + 16 x rounds tiles)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from sdrm_b200 import _lib
from sdrm_b200.engine import SamplerEngine
from sdrm_b200.train_SDRM import _resolve_schedule

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
w = bench.WORKLOADS["cfg5"]
dev = torch.device("cuda", 0)
diff, vae = bench.build_models(w, dev)
sched = _resolve_schedule(diff, w["T"], dev)


def engine(cluster, grid_limit=0):
    e = SamplerEngine(dev)
    e.pack_denoiser(diff, sched, w["nd"])
    e.pack_decoder(vae)
    e.set_option(_lib.OPT_CLUSTER, cluster)
    if grid_limit:
        e.set_option(_lib.OPT_GRID_LIMIT, grid_limit)
    return e


def timed(fn, reps=3):
    ms = []
    for i in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(i)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return ms


e2, e4, e2s = engine(2), engine(4), engine(2, 16)
n2, n4, ns = 148 * 128 * rounds, 132 * 128 * rounds, 16 * 128 * rounds
out = torch.empty((n2 + 16 * 128 * rounds, w["I"]), dtype=torch.float32, device=dev)
side = torch.cuda.Stream(dev)

ms = timed(lambda i: e2.sample(n2, seed=5 + i, out=out))
print(f"HYB cs2 alone: {n2 // 128} tiles ms {[round(m, 1) for m in ms]} tiles/ms {n2 / 128 / min(ms):.3f}", flush=True)
ms = timed(lambda i: e4.sample(n4, seed=5 + i, out=out))
print(f"HYB cs4 alone: {n4 // 128} tiles ms {[round(m, 1) for m in ms]} tiles/ms {n4 / 128 / min(ms):.3f} cluster {e4.lib.sdrm_last_launch_count(e4.handle)}", flush=True)
ms = timed(lambda i: e2s.sample(ns, seed=5 + i, out=out))
print(f"HYB cs2 on 16 CTAs alone: {ns // 128} tiles ms {[round(m, 1) for m in ms]}", flush=True)


def both(i):
    cur = torch.cuda.current_stream(dev)
    e4.sample(n4, seed=5 + i, out=out)
    side.wait_stream(cur) if False else None
    with torch.cuda.stream(side):
        e2s.sample(ns, seed=50 + i, row_offset=n4, out=out[n4:])
    cur.wait_stream(side)


ms = timed(both)
print(f"HYB cs4 on 132 + cs2 on 16: {(n4 + ns) // 128} tiles ms {[round(m, 1) for m in ms]} tiles/ms {(n4 + ns) / 128 / min(ms):.3f}", flush=True)
