"""Address-ordered stall profile: consecutive SASS instructions with the same source line are merged.
usage: ncu_addr.py X.csv [min_pct]   (X.csv from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
hdr = None; fname = ""; cur = None; ins = []
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r: continue
    nm = len(hdr) - 4
    if r[0].isdigit():
        cur = (fname, int(r[0]), ",".join(r[1:len(r) - nm - 2]).strip()[:70]); continue
    if r[0] == "" and len(r) > nm + 2 and r[2].startswith("0x"):
        m = dict(zip(hdr[4:], r[len(r) - nm:]))
        st = {k[6:]: int(v) for k, v in m.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
        ins.append((int(r[2], 16), cur, r[3].strip(), int(m["# Samples"]), int(m["Instructions Executed"]), st))
ins.sort()
tot = sum(i[3] for i in ins)
groups = []
for a, c, sass, s, n, st in ins:
    if groups and groups[-1][1] == c:
        g = groups[-1]; g[2] += s; g[3] = max(g[3], n)
        for k, v in st.items(): g[4][k] = g[4].get(k, 0) + v
        g[5] += 1
    else:
        groups.append([a, c, s, n, dict(st), 1])
base = ins[0][0]
acc = 0
for a, c, s, n, st, cnt in groups:
    acc += s
    if 100 * s / tot >= minp:
        top = ", ".join(f"{k}:{100*v/tot:.1f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"+{a-base:6x} {100*s/tot:5.1f}% (cum {100*acc/tot:5.1f}) x{n:<11d} {c[0]}:{c[1]:<4d} {c[2]}  [{top}]")
