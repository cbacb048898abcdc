import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
from amq_b200 import tp
from amq_b200.arch import LINEARS, ModelShape
from amq_b200.model import QuantDecoder
world, batch, fused = 4, 1, False
bad = 0
for trial in range(int(os.environ.get("TRIALS", "6"))):
    shape = ModelShape("tiny-gqa", 512, 1024, 8, 4, 2, 512, head_dim=64, qkv_bias=False)
    rs = np.random.RandomState(world)
    arch = {n: rs.choice([2, 3, 4], size=2).tolist() for n in LINEARS}
    full = QuantDecoder(shape, arch, batch=batch, max_seq=32, seed=5)
    grp = tp.LocalTPGroup(full, world, fused=fused)
    tok = torch.randint(0, shape.vocab, (batch,), device=full.dev)
    full.reset(); full.tokens.copy_(tok); grp.set_tokens(tok)
    for pos in range(8):
        for m in grp.ranks: m.tokens.copy_(full.tokens)
        torch.cuda.synchronize()
        if pos < 2: full.step_eager(); grp.step_eager()
        else: full.step(); grp.step()
        torch.cuda.synchronize()
        ref = full.logits
        diffs = [float((m.logits - grp.ranks[0].logits).abs().max()) for m in grp.ranks]
        rel = float((grp.ranks[0].logits - ref).abs().max() / ref.abs().max())
        if max(diffs) > 0 or rel > 2e-2:
            bad += 1
            print("trial", trial, "pos", pos, "rank diffs", diffs, "rel vs full", rel, "h diffs", [float((m.h - grp.ranks[0].h).abs().max()) for m in grp.ranks], flush=True)
            break
    print("trial", trial, "timeouts", grp.timeouts(), flush=True)
    del grp, full
print("bad trials:", bad)
