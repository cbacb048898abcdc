"""Per-launch-class timing of the model's own decode-step launches (CUDA graph per class, PDL on)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200.arch import MODELS, sample_arch
from amq_b200.model import QuantDecoder
from amq_b200._lib import lib, check, ptr, cur_stream
import ctypes

shape = MODELS["Llama-2-7b-hf"]
arch = sample_arch(shape, 3.0, seed=0)
B = int(os.environ.get("B", "1"))
nb = int(os.environ.get("BLOCKS", shape.n_block))
m = QuantDecoder(shape, arch, batch=B, max_seq=256, n_block=nb)
if os.environ.get("ALIAS_LAYERS") == "1":      # every layer reads layer 0's weights: the 82 MB stay in L2 (upper bound of a perfect prefetcher)
    from amq_b200.arch import LINEARS
    for Lr in m.layers[1:]:
        for name in LINEARS:
            Lr[name] = m.layers[0][name]
    m._build_problems()
m.pos.fill_(100)
L = lib()
L.amqb_set_pdl(1)

def timeit(fn, n_launch, iters=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(iters):
            g.replay()
        e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters / n_launch

for key, cnt in (("qkv", 3), ("o", 1), ("gu", 2), ("down", 1)):
    def fn(key=key, cnt=cnt):
        for P in m._plan:
            m._gemv(P[key], cnt)
    print(f"{key:5s}: {timeit(fn, len(m._plan)):6.2f} us per launch", flush=True)
def attn():
    for P in m._plan:
        Lr = P["L"]
        check(L.amqb_attn_decode(ptr(m.qkv), ptr(Lr["k_cache"]), ptr(Lr["v_cache"]), ptr(m.attn), ptr(m.pos), m.B, m.Hq, m.Hkv,
                                 m.D, m.max_seq, ctypes.c_float(shape.rope_theta), ptr(m.rope), cur_stream()))
print(f"attn : {timeit(attn, len(m._plan)):6.2f} us per launch (pos=100)")
def head():
    check(L.amqb_lm_head(ptr(m.lm_head), ptr(m.h), ptr(m.final_norm), ctypes.c_float(shape.rms_eps), ptr(m.logits), m.B,
                         shape.vocab, m.H, cur_stream()))
    check(L.amqb_argmax(ptr(m.logits), ptr(m.next_tokens), m.B, shape.vocab, cur_stream()))
print(f"lm_head+argmax: {timeit(head, 1):6.2f} us")
def full():
    m._step_launches()
t_full = timeit(full, 1)
print(f"full step: {t_full:8.1f} us  ({nb} blocks: {(t_full - 52.0) / nb:6.2f} us per layer excluding ~52 us of embed / lm_head / argmax)")
