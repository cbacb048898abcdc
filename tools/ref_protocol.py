#!/usr/bin/env python
"""The reference's own TPS protocol on config 2 (Llama-2-7B, AMQ avg 3.0 bits, batch 1):
benchmark_tps (/root/reference/amq/utils/speed.py:23-46) = wall clock around greedy generation of `gen`
tokens INCLUDING the prompt pass, TPS = gen / median over iterations; reference defaults prompt 64, gen 128.
Run with the prompt consumed by QuantDecoder.prefill (one pass, tcgen05 linears) and token by token, plus the
prompt pass alone at 63 and 511 rows.

    python tools/ref_protocol.py [--iters 5] [--out gpurun_out/ref_protocol.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200.arch import MODELS, sample_arch          # noqa: E402
from amq_b200.model import QuantDecoder                # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--prompt", type=int, default=64)
    ap.add_argument("--gen", type=int, default=128)
    ap.add_argument("--out", default="gpurun_out/ref_protocol.json")
    ap.add_argument("--model", default="Llama-2-7b-hf")
    ap.add_argument("--bits", type=float, default=3.0)
    a = ap.parse_args()
    shape = MODELS[a.model]
    arch = sample_arch(shape, a.bits, seed=0)
    m = QuantDecoder(shape, arch, batch=1, max_seq=max(768, a.prompt + a.gen + 8), seed=0)
    m.capture()
    ids = torch.randint(0, shape.vocab - 1, (1, a.prompt))        # speed.py:93 input_ids = randint(0, vocab-1, (b, seq))
    res = {"workload": f"{shape.name} AMQ avg {a.bits}, batch 1, prompt {a.prompt}, gen {a.gen} (speed.py:23-46 protocol)"}
    for mode in (True, False):
        m.generate(ids, 4, prefill=mode)                           # warm-up (lazy module loads)
        ts = []
        for _ in range(a.iters):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m.generate(ids, a.gen, prefill=mode)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        res["tps_prefill_pass" if mode else "tps_prompt_token_by_token"] = a.gen / float(np.median(ts))
    for T in (63, 511):
        p = torch.randint(0, shape.vocab - 1, (1, T), device=m.dev)
        m.reset(); m.prefill(p)
        ts = []
        for _ in range(a.iters):
            m.reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m.prefill(p); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        ls = shape.linear_shape                       # the last layer stops after q|k|v + cache append
        per_layer = sum(n * k for (n, k) in ls.values())
        tail = sum(ls[n][0] * ls[n][1] for n in ("self_attn.o_proj", "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"))
        flops = 2.0 * T * (per_layer * shape.n_block - tail)
        res[f"prefill_{T}_rows_ms"] = ms
        res[f"prefill_{T}_rows_linear_tflops"] = flops / (ms * 1e-3) / 1e12
    print(json.dumps(res))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
