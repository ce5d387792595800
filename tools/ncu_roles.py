"""Unique-address stall accounting per warp role of sdrm_layer_engine_kernel (role = address range found from known lines)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; fname = ""; cur = None; ins = {}; text = {}
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r: continue
    nm = len(hdr) - 4
    if r[0].isdigit():
        cur = (fname, int(r[0])); text[cur] = ",".join(r[1:len(r) - nm - 2]); continue
    if r[0] == "" and len(r) > nm + 2 and r[2].startswith("0x"):
        m = dict(zip(hdr[4:], r[len(r) - nm:]))
        st = {k[6:]: int(v) for k, v in m.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
        a = int(r[2], 16)
        e = ins.setdefault(a, dict(lines=set(), sass=r[3].strip(), s=int(m["# Samples"]), n=int(m["Instructions Executed"]), st=st))
        e["lines"].add(cur)
addrs = sorted(ins)
base = addrs[0]
import os
K = "layer_engine_kernel.cuh"
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sdrm_b200", "csrc", K)).read().splitlines()
def line_of(marker):
    return next(i + 1 for i, l in enumerate(src) if marker in l)
marks = [("producers", line_of("const bool is_w = (warp == W_WARP)")), ("mma", line_of("uint32_t stage = 0, sphase = 0, cc = 0")),
         ("epilogue", line_of("const int q = warp & 3;")), ("noise", line_of("const int full_groups = L >> 4;")),
         ("exit", line_of("nobody exits") - 3), ("end", len(src) + 1)]
def first_addr(lo, hi):
    for a in addrs:
        if any(l[0] == K and lo <= l[1] < hi for l in ins[a]["lines"]): return a
bounds = [("prologue", base)] + [(n, first_addr(lo, marks[i + 1][1])) for i, (n, lo) in enumerate(marks[:-1])]
bounds = [(n, a) for n, a in bounds if a is not None]
tot = sum(e["s"] for e in ins.values())
print("total unique samples", tot)
for i, (name, a0) in enumerate(bounds):
    a1 = bounds[i + 1][1] if i + 1 < len(bounds) else addrs[-1] + 16
    sel = [ins[a] for a in addrs if a0 <= a < a1]
    s = sum(e["s"] for e in sel)
    st = {}
    for e in sel:
        for k, v in e["st"].items(): st[k] = st.get(k, 0) + v
    def is_wait(e):
        return any(("mbar_try_wait" in text.get(l, "") or "nanosleep" in text.get(l, "") or "spins" in text.get(l, "") or
                    "mbarrier.try_wait" in text.get(l, "")) for l in e["lines"]) or "SYNCS.PHASECHK" in e["sass"] or "NANOSLEEP" in e["sass"]
    waits = sum(e["s"] for e in sel if is_wait(e))
    n_wait = sum(e["n"] for e in sel if is_wait(e))
    fence = sum(e["s"] for e in sel if "MEMBAR" in e["sass"] or "FENCE" in e["sass"])
    n = sum(e["n"] for e in sel)
    print(f"{name:10s} +{a0-base:6x}  samples {100*s/tot:5.1f}%  mbar-wait {100*waits/max(s,1):5.1f}% of role  fences {100*fence/max(s,1):5.1f}%  instr {n/1e9:.2f}G (wait loops {n_wait/1e9:.2f}G)  "
          + ", ".join(f"{k}:{100*v/max(s,1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:6]))
