"""Aggregate an ncu source page by source-line regions: samples and stall reasons per region.
usage: python tools/ncu_regions.py rep.ncu-rep file.cuh name:lo-hi [name:lo-hi ...]"""
import csv, subprocess, sys, collections
rep, fname = sys.argv[1], sys.argv[2]
regions = []
extra = []
args = sys.argv[3:]
if args and args[0].startswith("--skip="):
    extra = ["--launch-skip", args[0].split("=")[1], "--launch-count", "1"]; args = args[1:]
for a in args:
    n, r = a.split(":"); lo, hi = r.split("-"); regions.append((n, int(lo), int(hi)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + extra, capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi_ = [i for i, r in enumerate(rows) if r and r[0] == "Line No" and len(r) > 5]
hdr = rows[hi_[0]]
ci = {}
for i, h in enumerate(hdr): ci.setdefault(h, i)
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
cur = None
for r in rows:
    if not r or r[0] == "Line No" or len(r) < len(hdr): continue
    if r[0].isdigit(): cur = int(r[0])
    if r[2].startswith("0x") and cur is not None:
        try: s = int(r[ci["# Samples"]])
        except ValueError: continue
        name = "other"
        for n, lo, hi in regions:
            if lo <= cur <= hi: name = n; break
        a = agg[name]; a[0] += s; a[1] += int(r[ci["Instructions Executed"]] or 0)
        for c in stall:
            v = int(r[ci[c]] or 0)
            if v: a[2][c[6:]] += v
tot = sum(a[0] for a in agg.values()) or 1
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{n:22s} samples {a[0]:6d} {100*a[0]/tot:5.1f}%  inst {a[1]:8d}  " + ", ".join(f"{k}={v}" for k, v in a[2].most_common(6)))
