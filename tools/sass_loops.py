"""Static SASS loop report: innermost loops (backward branches) of a kernel with instruction mix.
usage: sass_loops.py lib.so kernel_substring"""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ins = []; on = False
for line in txt.splitlines():
    if "Function :" in line:
        on = pat in line
        continue
    if not on: continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2i = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:[!\w]+,\s*)?`?\(?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr2i: loops.append((addr2i[tgt], i))
# innermost: loops containing no other loop
inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
def mix(lo, hi):
    c = collections.Counter()
    for a, t in ins[lo:hi + 1]:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        c[op.split(".")[0]] += 1
    return c
print(f"{len(ins)} instructions, {len(loops)} loops, {len(inner)} innermost")
for lo, hi in sorted(loops, key=lambda l: l[0]):
    c = mix(lo, hi)
    n = hi - lo + 1
    tag = [k for k in ("LDTM", "UTCHMMA", "UTMALDG", "UBLKCP", "MUFU", "SYNCS", "NANOSLEEP") if c.get(k)]
    if n < 12 and not {"UTCHMMA", "UTMALDG", "UBLKCP"} & set(tag): continue
    kind = "inner" if (lo, hi) in inner else "outer"
    top = ", ".join(f"{k}:{v}" for k, v in c.most_common(9))
    print(f"{kind} +{ins[lo][0]:5x}..+{ins[hi][0]:5x} n={n:4d} LDL/STL={c.get('LDL',0)+c.get('STL',0):3d} {tag}  {top}")
