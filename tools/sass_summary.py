"""SASS opcode summary of the built library (profiles/sass_rNN.txt): proves which hardware paths each kernel uses.
usage: python tools/sass_summary.py [lib.so] > profiles/sass_r02.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "sdrm_b200/csrc/libsdrm_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMACMDFLUSH", "SYNCS", "HMMA", "LDGSTS", "LDSM", "MOVM", "CCTL", "MUFU", "FFMA2", "FADD2", "FMUL2",
        "ATOMS", "RED", "ATOMG", "STG", "LDG", "LDS", "STS", "SHFL", "LDL", "STL", "BAR", "NANOSLEEP", "ELECT"]
cur = None
ops = collections.OrderedDict()
n_ins = {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); ops[cur] = collections.Counter(); n_ins[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        n_ins[cur] += 1
        full = m.group(1)
        base = full.split(".")[0]
        ops[cur][base] += 1
        if base in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "CCTL", "LDG", "STG"):
            ops[cur][full] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(ops), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {lib}: instruction counts per kernel (static), sm_100a")
for (k, c), name in zip(ops.items(), demangle):
    short = re.sub(r"\(.*", "", name)[:100]
    keys = [w for w in WANT if c.get(w)]
    detail = sorted((k2, v) for k2, v in c.items() if "." in k2 and k2.split(".")[0] in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "CCTL"))
    print(f"\n{short}   [{n_ins[k]} instructions]")
    print("   " + "  ".join(f"{w}:{c[w]}" for w in keys))
    if detail:
        print("   " + "  ".join(f"{a}:{b}" for a, b in detail))
