"""Summarise an ncu launch list (csv with gpu__time_duration.sum per launch): time share per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    v = float(r[ci["Metric Value"]])
    if r[ci["Metric Unit"]] == "ns":
        v /= 1000
    agg[r[ci["Kernel Name"]][:70]].append(v)
tot = sum(sum(v) for v in agg.values())
print(f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v):9.1f} us {100*sum(v)/tot:5.1f}%  n={len(v):4d} avg={sum(v)/len(v):7.2f} min={min(v):6.2f} max={max(v):6.2f}  {k}")
