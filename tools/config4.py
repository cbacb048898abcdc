"""Config 4 (BASELINE.json): Qwen2-7B proxy sweep, quantize -> pack -> dequantize for every linear shape at 2 / 3 / 4 bits,
in both solver arithmetics (fp16 = the reference's CUDA branch, fp32 = its CPU branch, bit-exact)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amq_b200
from amq_b200.arch import MODELS

dev = torch.device("cuda")
shape = MODELS["Qwen2.5-7B"]
out = {"n_block": shape.n_block}
for mode, sd in (("fp16", torch.float16), ("fp32", torch.float32)):
    amq_b200.Quantizer.solver_dtype = sd
    total = {2: 0.0, 3: 0.0, 4: 0.0}
    per = {}
    for name, (N, K) in shape.linear_shape.items():
        torch.manual_seed(0)
        W = (torch.randn(N, K, device=dev) * 0.02).half()
        for bits in (2, 3, 4):
            cfg = amq_b200.BaseQuantizeConfig(nbits=bits, group_size=128)["weight_quant_params"]
            def run():
                W_q, meta = amq_b200.Quantizer.quantize(W, device="cuda", compute_dtype=torch.float16, **cfg)
                meta16 = dict(meta, scale=meta["scale"].half(), zero=meta["zero"].half())
                return amq_b200.Quantizer.dequantize(W_q, meta16)
            run(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            alg = 2 * N * K + N * K * bits // 8 + (N * K // 128) * 4
            alg += N * K * bits // 8 + (N * K // 128) * 4 + 2 * N * K
            per[f"{name}/{bits}bit"] = {"ms": round(ms, 4), "algorithmic_GBps": round(alg / ms / 1e6, 1)}
            total[bits] += ms
    out[mode] = {f"sweep_ms_all_linears_{b}bit": round(total[b] * shape.n_block, 2) for b in (2, 3, 4)}
    out[mode]["per_layer"] = per
    print(mode, {k: v for k, v in out[mode].items() if k != "per_layer"}, flush=True)
amq_b200.Quantizer.solver_dtype = None
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r02_config4.json", "w"), indent=1)
