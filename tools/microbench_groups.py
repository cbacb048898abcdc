"""Microbenchmark of the decode GEMV launches exactly as the model issues them: q|k|v and gate|up grouped
with the RMSNorm prologue, o_proj with residual, down_proj with SiLU*up prologue + residual (Llama-2-7B
shapes), rotating over enough layer copies to defeat L2."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200 import ops
from amq_b200._lib import PRO_NONE, PRO_RMSNORM, PRO_SILU_MUL

dev = torch.device("cuda")
H, I = 4096, 11008
COPIES = 12

KEEP = []


def nat(bits, N, K):
    t = torch.randint(0, 256, (ops.native_bytes(bits, N, K),), dtype=torch.uint8, device=dev)
    KEEP.append(t)          # problems hold raw pointers only
    return t

def bench(name, make_group, alg_bytes, pdl=True, iters=20):
    groups = [make_group(i) for i in range(COPIES)]
    ws = ops.workspace(dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for g_ in groups[:2]:
            ops.gemv_grouped(g_, ws, pdl=False)
        s.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for g_ in groups:
                ops.gemv_grouped(g_, ws, pdl=pdl)
        gr.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(iters):
            gr.replay()
        e1.record(s); s.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (iters * COPIES)
    KEEP.clear()
    print(json.dumps({"case": name, "us": round(us, 2), "GBps": round(alg_bytes / us / 1e3, 1)}), flush=True)

h = torch.randn(1, H, device=dev).half()
gamma = torch.ones(H, device=dev).half()
qkv = torch.zeros(1, 3 * H, device=dev).half()
attn = torch.randn(1, H, device=dev).half()
gu = torch.randn(1, 2 * I, device=dev).half()
hres = torch.randn(1, H, device=dev).half()

def bytes_of(bits, N, K):
    return N * K * bits // 8 + (K // 128) * N * 4

for bq, bk, bv in [(3, 3, 3), (2, 3, 4), (4, 4, 4), (2, 2, 2)]:
    def mk(i, bq=bq, bk=bk, bv=bv):
        ps = []
        for j, b in enumerate((bq, bk, bv)):
            p = ops.make_problem(b, nat(b, H, H), h, qkv, H, H, prologue=PRO_RMSNORM, gamma=gamma, eps=1e-5, ldy=3 * H)
            p.y = qkv.data_ptr() + 2 * j * H
            ps.append(p)
        return ps
    bench(f"qkv rms bits={bq}{bk}{bv}", mk, sum(bytes_of(b, H, H) for b in (bq, bk, bv)))
    def mk1(i, b=bq):
        return [ops.make_problem(b, nat(b, H, H), h, qkv, H, H, prologue=PRO_RMSNORM, gamma=gamma, eps=1e-5, ldy=3 * H)]
    bench(f"q only rms bits={bq}", mk1, bytes_of(bq, H, H))
for bg, bu in [(3, 3), (2, 4)]:
    def mk(i, bg=bg, bu=bu):
        ps = []
        for j, b in enumerate((bg, bu)):
            p = ops.make_problem(b, nat(b, I, H), h, gu, I, H, prologue=PRO_RMSNORM, gamma=gamma, eps=1e-5, ldy=2 * I)
            p.y = gu.data_ptr() + 2 * j * I
            ps.append(p)
        return ps
    bench(f"gate|up rms bits={bg}{bu}", mk, sum(bytes_of(b, I, H) for b in (bg, bu)))
for b in (2, 3, 4):
    bench(f"o_proj residual bits={b}", lambda i, b=b: [ops.make_problem(b, nat(b, H, H), attn, hres, H, H, residual=hres)], bytes_of(b, H, H))
    bench(f"down silu residual bits={b}", lambda i, b=b: [ops.make_problem(b, nat(b, H, I), gu, hres, H, I, residual=hres, prologue=PRO_SILU_MUL, ldx=2 * I)], bytes_of(b, H, I))
