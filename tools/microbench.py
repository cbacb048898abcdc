"""Kernel-level microbenchmark of the decode GEMV (run on the GPU box).

Rotates over enough distinct weight copies to exceed L2 (126 MB) so the numbers are HBM numbers,
replays a CUDA graph of the launches, and times with CUDA events on the launching stream.
Prints one JSON line per case: achieved algorithmic GB/s = bytes(N,K,b,M) / time.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200 import ops  # noqa: E402

G = 128


def alg_bytes(N, K, b, M):
    return N * K * b // 8 + (K // G) * N * 4 + 2 * M * K + 2 * M * N


def bench_case(N, K, bits, M, pool_mb=384, iters=20, pdl=False, simt=False):
    dev = torch.device("cuda")
    nb = ops.native_bytes(bits, N, K)
    copies = max(2, int(pool_mb * 2 ** 20 / nb))
    gen = torch.Generator(device="cuda").manual_seed(1)
    pool = [torch.randint(0, 256, (nb,), dtype=torch.uint8, device=dev, generator=gen) for _ in range(copies)]
    # sane meta: overwrite by packing a real synthetic layer into copy 0.. (values irrelevant for timing)
    x = torch.randn(M, K, device=dev).half()
    y = torch.empty(M, N, device=dev, dtype=torch.float16)
    ws = ops.workspace(dev, N, K, M)
    probs = [ops.make_problem(bits, w, x, y, N, K) for w in pool]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for p in probs[:3]:
            ops.gemv_grouped([p], ws, pdl=False)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for p in probs:
                ops.gemv_grouped([p], ws, pdl=pdl)
        for _ in range(3):
            g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(iters):
            g.replay()
        e1.record(s)
        s.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (iters * copies)
    gbs = alg_bytes(N, K, bits, M) / us / 1e3
    return {"N": N, "K": K, "bits": bits, "M": M, "pdl": pdl, "copies": copies, "us": round(us, 3),
            "GBps": round(gbs, 1), "magic": os.environ.get("AMQB_GEMV_MAGIC", "0")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--out", default="gpurun_out/microbench.jsonl")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    shapes = [(4096, 4096), (11008, 4096), (4096, 11008)]
    if a.tiny:
        shapes = [(4096, 4096), (11008, 4096)]
    if not a.quick and not a.tiny:
        shapes += [(8192, 8192), (28672, 8192), (8192, 28672), (1024, 4096)]
    rows = []
    for (N, K) in shapes:
        for bits in (2, 3, 4):
            for M in ((1,) if (a.quick or a.tiny) else (1, 4, 16)):
                for pdl in ((True,) if a.tiny else (False, True)):
                    r = bench_case(N, K, bits, M, pdl=pdl)
                    rows.append(r)
                    print(json.dumps(r), flush=True)
    with open(a.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
