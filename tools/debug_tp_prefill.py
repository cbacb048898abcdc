"""Per-layer K/V error of the emulated tensor-parallel prompt pass against the unsharded one (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from amq_b200 import tp
from amq_b200.arch import LINEARS, ModelShape
from amq_b200.model import QuantDecoder


def run(world, batch, prompt, bias):
    shape = ModelShape("tiny-gqa", 512, 1024, 8, 4, 2, 512, head_dim=64, qkv_bias=bias)
    rs = np.random.RandomState(world)
    arch = {n: rs.choice([2, 3, 4], size=2).tolist() for n in LINEARS}
    full = QuantDecoder(shape, arch, batch=batch, max_seq=64, seed=5)
    grp = tp.LocalTPGroup(full, world, fused=True)
    ids = torch.randint(0, shape.vocab, (batch, prompt), device=full.dev)
    P = prompt - 1
    full.reset(); grp.set_tokens(ids[:, 0])
    full.prefill(ids[:, :P]); grp.prefill(ids[:, :P])
    torch.cuda.synchronize()
    hk = full.Hkv // world
    out = []
    for r, m in enumerate(grp.ranks):
        for li, (Lf, Lm) in enumerate(zip(full.layers, m.layers)):
            for c in ("k_cache", "v_cache"):
                want, got = Lf[c][:, r * hk:(r + 1) * hk, :P].float(), Lm[c][:, :, :P].float()
                out.append((r, li, c[0], round(float((want - got).abs().max() / want.abs().max()), 4)))
    print(f"world={world} batch={batch} prompt={prompt} bias={bias} tc={'off' if os.environ.get('AMQB_NO_TCGEN05') else 'on'} "
          f"timeouts={grp.timeouts()}:", [o for o in out if o[3] > 1e-2] or "all ok", flush=True)


for cfg in [(4, 2, 30, True), (4, 2, 30, False), (4, 1, 30, False), (4, 1, 59, False), (4, 2, 6, True), (2, 2, 30, True)]:
    run(*cfg)
os.environ["AMQB_NO_TCGEN05"] = "1"
run(4, 2, 30, True)
