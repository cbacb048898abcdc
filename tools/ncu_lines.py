"""Summarise an ncu report: headline metrics + top source lines by stall samples.
usage: python tools/ncu_lines.py rep.ncu-rep [n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "sm__cycles_active.avg"]
for i, h in enumerate(hdr):
    if h in want or h == "dram__bytes_read.sum.per_second":
        print(f"{h} [{units[i]}]: {[r[i] for r in data]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Line No" and len(r) > 5]
hdr = rows[hdrs[0]]
ci = {}
for i, h in enumerate(hdr):
    ci.setdefault(h, i)
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
cur = None
agg = {}
kernel_idx = 0
for idx, r in enumerate(rows):
    if not r:
        continue
    if r[0] == "Line No":
        continue
    if len(r) < len(hdr):
        continue
    if r[0].isdigit():
        cur = (int(r[0]), r[1].strip())
    if r[2].startswith("0x") and cur:
        try:
            s = int(r[ci["# Samples"]])
        except ValueError:
            continue
        a = agg.setdefault(cur, [0, 0, {}])
        a[0] += s
        a[1] += int(r[ci["Instructions Executed"]] or 0)
        for c in stall:
            v = int(r[ci[c]] or 0)
            if v:
                a[2][c] = a[2].get(c, 0) + v
tot = sum(a[0] for a in agg.values()) or 1
print("total samples", tot, " total inst", sum(a[1] for a in agg.values()))
for (ln, s_), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    top = sorted(a[2].items(), key=lambda kv: -kv[1])[:3]
    print(f"{a[0]:6d} {100*a[0]/tot:5.1f}% L{ln:<4d} inst={a[1]:<8d} {s_[:78]}  {[(k[6:], v) for k, v in top]}")
print("---- top lines by instructions executed")
for (ln, s_), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
    print(f"inst={a[1]:<8d} samples={a[0]:<5d} L{ln:<4d} {s_[:90]}")
