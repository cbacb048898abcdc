"""Measurements for BASELINE.json configs 3 and 4 (single GPU); config 2 is bench.py, config 5 is
bench.py --workload llama70b-tp.  Writes one JSON document to gpurun_out/configs.json."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200 import ops
from amq_b200.arch import MODELS, sample_arch, get_bits_usage, LINEARS
from amq_b200.model import QuantDecoder, synthetic_native
import amq_b200

out = {}
dev = torch.device("cuda")
PEAK_HBM, PEAK_TF = 6549.4, 1664.0
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    d = json.load(open(p)); PEAK_HBM, PEAK_TF = d["hbm_gbs"], d["bf16_tflops"]

def time_steps(model, steps=64, warm=8):
    model.reset(); model.tokens.fill_(1)
    for _ in range(warm): model.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): model.step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

# ---- config 3: Mistral-7B, avg 2.5 bits, decode batch 1 / 4 / 16 + 512-token prefill of the linears
shape = MODELS["Mistral-7B-v0.3"]
arch = sample_arch(shape, 2.5, seed=0)
c3 = {"bits_usage": get_bits_usage({"linear": arch}, shape.config()), "decode": {}}
for B in (1, 4, 16):
    m = QuantDecoder(shape, arch, batch=B, max_seq=256)
    ms = time_steps(m)
    by = m.algorithmic_bytes_per_token()["total"]
    c3["decode"][f"batch{B}"] = {"ms_per_step": ms, "tok_per_s": B * 1e3 / ms, "frac_of_hbm_roofline": by / (PEAK_HBM * 1e9) / (ms * 1e-3)}
    print("config3 decode", B, c3["decode"][f"batch{B}"], flush=True)
    del m; torch.cuda.empty_cache()
# prefill: every quantized linear of the model at M = 512 through amqb_gemm_tc (tcgen05), attention not included
M = 512
gen = torch.Generator(device=dev).manual_seed(0)
tot_ms, tot_flop = 0.0, 0.0
per_shape = {}
for name, (N, K) in shape.linear_shape.items():
    for bits in (2, 3, 4):
        cnt = sum(1 for b in arch[name] if b == bits)
        if cnt == 0: continue
        w = synthetic_native(bits, N, K, dev, gen)
        x = torch.randn(M, K, device=dev).half()
        for _ in range(3): ops.gemm_tc(bits, w, x, N, K)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.gemm_tc(bits, w, x, N, K)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        per_shape[f"{name}/{bits}bit"] = {"us": ms * 1e3, "tflops": 2 * M * N * K / ms / 1e9, "count": cnt}
        tot_ms += ms * cnt; tot_flop += 2.0 * M * N * K * cnt
c3["prefill512_linears"] = {"ms_all_224_linears": tot_ms, "tflops": tot_flop / tot_ms / 1e9, "frac_of_bf16_peak": tot_flop / tot_ms / 1e9 / PEAK_TF,
                            "tokens_per_s_linears_only": M * 1e3 / tot_ms, "per_shape": per_shape}
print("config3 prefill", {k: v for k, v in c3["prefill512_linears"].items() if k != "per_shape"}, flush=True)
# the same 224 linears as the prompt pass launches them: q|k|v and gate|up grouped (one launch each), o_proj and down_proj
# single, layer by layer with the model's own bit widths, captured in one CUDA graph
nat = {}
def weight(name, bits):
    if (name, bits) not in nat:
        N, K = shape.linear_shape[name]
        nat[(name, bits)] = synthetic_native(bits, N, K, dev, gen)
    return nat[(name, bits)]
H, I = shape.hidden, shape.inter
xh = torch.randn(M, H, device=dev).half(); xa = torch.randn(M, shape.n_heads * shape.head_dim, device=dev).half(); xi = torch.randn(M, I, device=dev).half()
def layer(li):
    b = {n: int(arch[n][li % len(arch[n])]) for n in LINEARS}
    ops.linear_forward_grouped([(b[n], weight(n, b[n]), shape.linear_shape[n][0], None) for n in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj")], xh, H)
    n = "self_attn.o_proj"; ops.gemm_tc(b[n], weight(n, b[n]), xa, *shape.linear_shape[n])
    ops.linear_forward_grouped([(b[n], weight(n, b[n]), shape.linear_shape[n][0], None) for n in ("mlp.gate_proj", "mlp.up_proj")], xh, H)
    n = "mlp.down_proj"; ops.gemm_tc(b[n], weight(n, b[n]), xi, *shape.linear_shape[n])
for li in range(shape.n_block): layer(li)
torch.cuda.synchronize()
st_ = torch.cuda.Stream()
with torch.cuda.stream(st_):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st_):
        for li in range(shape.n_block): layer(li)
    gr.replay(); st_.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st_)
    for _ in range(5): gr.replay()
    e1.record(st_); st_.synchronize()
gms = e0.elapsed_time(e1) / 5
c3["prefill512_grouped"] = {"ms_all_224_linears": gms, "tflops": tot_flop / gms / 1e9, "frac_of_bf16_peak": tot_flop / gms / 1e9 / PEAK_TF,
                            "how": "q|k|v and gate|up as grouped launches, model order, one CUDA graph"}
print("config3 prefill grouped", c3["prefill512_grouped"], flush=True)
out["config3_mistral7b_2.5bit"] = c3

# ---- config 4: Qwen2-7B proxy sweep: quantize -> pack -> dequantize for every linear shape at 2/3/4 bits
shape = MODELS["Qwen2.5-7B"]
c4 = {"per_layer": {}, "n_block": shape.n_block}
total = {2: 0.0, 3: 0.0, 4: 0.0}
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
        alg = 2 * N * K + N * K * bits // 8 + (N * K // 128) * 4          # quantize+pack: read fp16, write codes+meta
        alg += N * K * bits // 8 + (N * K // 128) * 4 + 2 * N * K          # dequant: read codes+meta, write fp16
        c4["per_layer"][f"{name}/{bits}bit"] = {"ms": ms, "algorithmic_GBps": alg / ms / 1e6}
        total[bits] += ms
for bits in (2, 3, 4):
    c4[f"sweep_ms_all_linears_{bits}bit"] = total[bits] * shape.n_block
    c4[f"layers_per_s_{bits}bit"] = 7 * 1e3 / total[bits]
print("config4", {k: v for k, v in c4.items() if k != "per_layer"}, flush=True)
out["config4_qwen2_proxy_sweep"] = c4
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
