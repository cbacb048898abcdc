"""Summarise an ncu source page: top source lines by warp-stall samples.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda > src.csv ; python tools/ncu_top.py src.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No" and len(r) > 5)
hdr = rows[hdr_i]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
out = []
tot = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    s = int(r[ci["# Samples"]] or 0)
    tot += s
    if s:
        st = sorted(((int(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:3]
        out.append((s, int(r[0]), r[1].strip()[:90], int(r[ci["Instructions Executed"]] or 0), st))
out.sort(reverse=True)
print("total samples", tot)
for s, ln, src, ie, st in out[:n]:
    print(f"{s:6d} {100*s/tot:5.1f}%  L{ln:<4d} inst={ie:<8d} {src}   {[(c[6:], v) for v, c in st if v]}")
