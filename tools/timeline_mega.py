"""Phase timeline of the persistent decode kernel (layer 1), AMQB_TIMELINE build.  Stamps are SM clocks; each CTA's
stamps are taken relative to its own 'q|k|v wait done' (the grid barrier releases all CTAs within a fraction of a us)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200.arch import MODELS, sample_arch
from amq_b200.model import QuantDecoder
from amq_b200 import _lib
shape = MODELS["Llama-2-7b-hf"]
arch = sample_arch(shape, 3.0, seed=0)
if os.environ.get("UNIFORM"):
    from amq_b200.arch import LINEARS
    arch = {n: [int(os.environ["UNIFORM"])] * shape.n_block for n in LINEARS}
m = QuantDecoder(shape, arch, batch=1, max_seq=256)
m.pos.fill_(100)
for _ in range(3): m.step_eager()
torch.cuda.synchronize()
dbg = torch.zeros(148 * 5 * 8, dtype=torch.int64, device="cuda")
_lib.lib().amqb_debug_set_timeline(ctypes.c_void_p(dbg.data_ptr()))
m.step_eager(); torch.cuda.synchronize()
_lib.lib().amqb_debug_set_timeline(None)
d = dbg.cpu().view(148, 5, 8).double()
t0 = d[:, 0, 1].clone().view(148, 1, 1)
d = torch.where(d > 0, (d - t0) / 1.965e3, torch.full_like(d, float("nan")))
names = ["qkv", "attn", "o", "gu", "down"]
ev = {0: "phase entered", 1: "wait done", 2: "x' built / qkv loaded", 3: "attn loop done", 6: "attn cta barrier", 4: "last deposit / attn arrived", 5: "reducer arrived"}
for t in range(5):
    for e, nm in ev.items():
        c = d[:, t, e]; c = c[~torch.isnan(c)]
        if c.numel(): print(f"{names[t]:5s} {nm:28s} n={c.numel():3d} min {c.min():7.2f} med {c.median():7.2f} max {c.max():7.2f} us")
