"""Stall samples per SASS instruction inside an address range of the kernel (offsets relative to the first instruction).
usage: ncu_range.py source.csv 0xLO 0xHI [top]   (source.csv = ncu --page source --print-source sass,cuda --csv)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
hdr = None; ins = {}
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r: continue
    nm = len(hdr) - 4
    if r[0] == "" and len(r) > nm + 2 and r[2].startswith("0x"):
        m = dict(zip(hdr[4:], r[len(r) - nm:]))
        a = int(r[2], 16)
        if a in ins: continue
        st = {k[6:]: int(v) for k, v in m.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
        ins[a] = (r[3].strip(), int(m["# Samples"]), int(m["Instructions Executed"]), st)
base = min(ins)
sel = [(a - base, *ins[a]) for a in sorted(ins) if lo <= a - base <= hi]
tot = sum(v[1] for v in ins.values()); s = sum(x[2] for x in sel); n = sum(x[3] for x in sel)
print(f"range +{lo:x}..+{hi:x}: {len(sel)} instr, samples {s} = {100*s/tot:.1f}% of kernel, executed {n/1e6:.1f} M warp instr (max per instr {max(x[3] for x in sel)/1e6:.2f} M)")
agg = {}
for x in sel:
    for k, v in x[4].items(): agg[k] = agg.get(k, 0) + v
print("stall mix:", ", ".join(f"{k}:{100*v/max(s,1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for x in sorted(sel, key=lambda x: -x[2])[:top]:
    st = ", ".join(f"{k}:{v}" for k, v in sorted(x[4].items(), key=lambda kv: -kv[1])[:3])
    print(f"  +{x[0]:5x} {100*x[2]/max(s,1):5.1f}%  exec {x[3]/1e6:6.2f}M  {x[1][:70]:70s} [{st}]")
