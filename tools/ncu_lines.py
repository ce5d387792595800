"""Per-source-line stall samples / executed instructions from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
out = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r or r[0] == "" or not r[0].isdigit():
        continue
    nm = len(hdr) - 4   # source text may contain unescaped quotes / commas: take the metric columns from the right
    src = ",".join(r[1:len(r) - nm - 2])
    r = [r[0], src, "-", "-"] + r[len(r) - nm:]
    d = dict(zip(hdr[4:], r[4:]))
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
    out.append((int(d["# Samples"]), int(d["Instructions Executed"]), fname, int(r[0]), r[1].strip()[:90], stalls))
tot_s = sum(o[0] for o in out); tot_i = sum(o[1] for o in out)
print(f"total samples {tot_s}, total warp instructions {tot_i}")
print("== by samples")
for o in sorted(out, key=lambda o: -o[0])[:top]:
    st = ", ".join(f"{k}:{v}" for k, v in sorted(o[5].items(), key=lambda kv: -kv[1])[:4])
    print(f"{100*o[0]/tot_s:5.1f}% smp {100*o[1]/tot_i:5.1f}% ins  {o[2]}:{o[3]:<4d} {o[4]}\n        [{st}]")
print("== by instructions")
for o in sorted(out, key=lambda o: -o[1])[:top]:
    print(f"{100*o[1]/tot_i:5.1f}% ins {100*o[0]/tot_s:5.1f}% smp  {o[2]}:{o[3]:<4d} {o[4]}")
