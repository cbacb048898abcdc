"""Decode step of a tiny model in three modes (eager, graph without PDL, graph with PDL) against the torch fp32
reference.  Catches ordering bugs that only show under PDL."""
import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from test_gpu_model import _ref_step
from amq_b200.arch import ModelShape, LINEARS
from amq_b200.model import QuantDecoder
shape = ModelShape("tiny-llama", 256, 512, 4, 4, 2, 512, head_dim=64)
rs = np.random.RandomState(0)
arch = {n: rs.choice([2, 3, 4], size=2).tolist() for n in LINEARS}
for mode in ("eager", "graph_nopdl", "graph"):
    m = QuantDecoder(shape, arch, batch=1, max_seq=32, seed=1, pdl=(mode == "graph"))
    kc = [torch.zeros(1, m.Hkv, 32, m.D, device=m.dev) for _ in m.layers]
    vc = [torch.zeros(1, m.Hkv, 32, m.D, device=m.dev) for _ in m.layers]
    tok = torch.randint(0, shape.vocab, (1,), device=m.dev)
    m.reset(); m.tokens.copy_(tok)
    for pos in range(4):
        cur = m.tokens.clone()
        ref = _ref_step(m, cur, pos, kc, vc)
        if mode == "eager": m.step_eager()
        else: m.step()
        torch.cuda.synchronize()
        rel = (m.logits - ref).abs().max() / ref.abs().max()
        print(mode, pos, float(rel), flush=True)
