"""Summarise an `ncu --set full` capture of consecutive decode-GEMV launches taken inside bench.py into
profiles/r01_ncu_gemv_traffic.json (DRAM bytes per launch next to the kernel's duration and pipe utilisation).
usage: python tools/ncu_traffic.py rep.ncu-rep out.json"""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return float(r[ci[k]].replace(",", ""))
    except Exception: return None
def scaled(r, k, target):          # ncu prints byte / time metrics with a unit column: normalise
    v = num(r, k)
    if v is None: return None
    u = units[ci[k]].lower()
    mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return v * mult
launches = []
for r in data:
    if len(r) < len(hdr): continue
    launches.append({
        "kernel": r[ci["Kernel Name"]][:60], "grid": r[ci["Grid Size"]] if "Grid Size" in ci else None,
        "dur_us": scaled(r, "gpu__time_duration.sum", "us"),
        "dram_read_B": scaled(r, "dram__bytes_read.sum", "byte"), "dram_write_B": scaled(r, "dram__bytes_write.sum", "byte"),
        "dram_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "imma_pct": num(r, "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active") if "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active" in ci else None,
        "regs": r[ci["launch__registers_per_thread"]], "inst": num(r, "smsp__inst_executed.sum"),
    })
tot = [l["dram_read_B"] + (l["dram_write_B"] or 0) for l in launches if l["dram_read_B"] is not None]
doc = {"source": "ncu --set full --clock-control none, consecutive decode-GEMV launches of one decoder layer (qkv, o_proj, gate|up, "
                 "down_proj) inside bench.py (graph replay)", "launches": launches,
       "traffic_bytes_per_launch_avg": sum(tot) / len(tot) if tot else None}
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(doc, indent=1)[:1500])
