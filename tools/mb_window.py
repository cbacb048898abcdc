"""Producer in-flight window (AMQB_WINDOW_KB) on long weight streams: kernel-level GB/s of the large shapes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from microbench import bench_case, alg_bytes
for wkb in ("0", "64", "96", "128", "160"):
    os.environ["AMQB_WINDOW_KB"] = wkb
    row = {"window_kb": wkb}
    for (N, K) in [(28672, 8192), (8192, 28672), (11008, 4096), (4096, 4096)]:
        for bits in (3, 4):
            row[f"{N}x{K}/{bits}b"] = bench_case(N, K, bits, 1, pdl=True)["GBps"]
    print(json.dumps(row), flush=True)
