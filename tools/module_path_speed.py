"""Cost of staying drop-in at the module level (VERDICT r1 next #8c): Llama-2-7B-shaped HF decoder (random init) whose 224
linears are this library's GPTQLinear (2/3-bit) / FT_QuantLinear (4-bit) modules, swapped in by setattr like
amq/amq_speed_benchmark.py:231-251 (synthetic packed buffers, AMQ avg 3.0 bits), decoded token by token through HF's own
forward with its KV cache — 224 module.forward calls per token — next to QuantDecoder's captured step on the same shapes."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import transformers
import amq_b200
from amq_b200.arch import MODELS, sample_arch

dev = "cuda"
shape = MODELS["Llama-2-7b-hf"]
arch = sample_arch(shape, 3.0, seed=0)
nb = int(os.environ.get("BLOCKS", shape.n_block))
cfg = transformers.LlamaConfig(hidden_size=shape.hidden, intermediate_size=shape.inter, num_hidden_layers=nb,
                               num_attention_heads=shape.n_heads, num_key_value_heads=shape.n_kv_heads, vocab_size=shape.vocab,
                               max_position_embeddings=512, rms_norm_eps=shape.rms_eps, tie_word_embeddings=False)
with torch.device(dev):
    model = transformers.LlamaForCausalLM(cfg).half().eval()
gen = torch.Generator(device=dev).manual_seed(0)
for li, layer in enumerate(model.model.layers):
    for name, bl in arch.items():
        mod, lin = name.split(".")
        src = getattr(getattr(layer, mod), lin)
        K, N = src.in_features, src.out_features
        bits = int(bl[li % len(bl)])
        lo, hi = {2: (0.024, 0.054), 3: (0.010, 0.023), 4: (0.0047, 0.011)}[bits]
        sc = torch.empty(K // 128, N, device=dev).uniform_(lo, hi, generator=gen).half()
        zs = (torch.empty(K // 128, N, device=dev).uniform_(0.5, 2 ** bits - 1.5, generator=gen).half() * sc)
        if bits == 4:
            m = amq_b200.FT_QuantLinear(4, K, N, bias=False, dtype=torch.float16, group_size=128, name=name).to(dev)
            m.qweight = torch.randint(-2 ** 15, 2 ** 15 - 1, m.qweight.shape, dtype=torch.int16, device=dev, generator=gen)
            m.scales, m.scaled_zeros = sc, -zs
        else:
            m = amq_b200.GPTQLinear(bits, 128, K, N, bias=False).to(dev)
            m.qweight = torch.randint(-2 ** 31, 2 ** 31 - 1, m.qweight.shape, dtype=torch.int32, device=dev, generator=gen)
            m.scales, m.zeros = sc.float(), zs.float()
        delattr(getattr(layer, mod), lin)
        setattr(getattr(layer, mod), lin, m)
torch.cuda.synchronize()
ids = torch.randint(0, shape.vocab - 1, (1, 64), device=dev)
steps = int(os.environ.get("STEPS", "64"))
# torch 2.11's default SDPA choice on B200 (cuDNN) builds a plan per new (q_len, kv_len): 3.7 ms of HOST time per call when the
# cache grows by one position per token (tools/profile_hf_step.py).  The reference's protocol does not pin a backend, so both
# are reported: the default, and SDPA restricted to the flash / memory-efficient / math kernels.
from torch.nn.attention import sdpa_kernel, SDPBackend
res = {}
def hf_tok_s():
    with torch.inference_mode():
        out = model(ids, use_cache=True)
        past, tok = out.past_key_values, out.logits[:, -1:].argmax(-1)
        for _ in range(8):
            out = model(tok, past_key_values=past, use_cache=True); past, tok = out.past_key_values, out.logits[:, -1:].argmax(-1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = model(tok, past_key_values=past, use_cache=True); past, tok = out.past_key_values, out.logits[:, -1:].argmax(-1)
        torch.cuda.synchronize()
        return steps / (time.perf_counter() - t0)
with sdpa_kernel([SDPBackend.FLASH_ATTENTION, SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH]):
    res["module_path_tok_s_sdpa_no_cudnn"] = hf_tok_s()
from amq_b200.hf import GraphedHFDecoder
dec = GraphedHFDecoder(model, max_cache_len=256)
dec.prefill(ids)
for _ in range(8): dec.step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(steps): dec.step()
torch.cuda.synchronize()
res["module_path_graphed_tok_s"] = steps / (time.perf_counter() - t0)
del dec
with torch.inference_mode():
    out = model(ids, use_cache=True)
    past, tok = out.past_key_values, out.logits[:, -1:].argmax(-1)
    for _ in range(8):
        out = model(tok, past_key_values=past, use_cache=True); past, tok = out.past_key_values, out.logits[:, -1:].argmax(-1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = model(tok, past_key_values=past, use_cache=True); past, tok = out.past_key_values, out.logits[:, -1:].argmax(-1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
res.update({"module_path_tok_s": steps / dt, "ms_per_token": dt / steps * 1e3, "blocks": nb, "forward_calls_per_token": 7 * nb,
       "how": "HF LlamaForCausalLM.forward with DynamicCache, eager, 7 quantized-linear module calls per block"})
del model, past, out
torch.cuda.empty_cache()
from amq_b200.model import QuantDecoder
import dataclasses
qd = QuantDecoder(dataclasses.replace(shape, n_block=nb), arch, batch=1, max_seq=256)
qd.capture(); qd.reset()
for _ in range(8): qd.step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(steps): qd.step()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
res["quantdecoder_tok_s"] = steps / dt
print(json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r02_module_path_speed.json", "w"), indent=1)
