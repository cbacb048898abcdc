"""Per-problem timeline of a grouped q|k|v launch as the model issues it (stamps of consumer warp 0)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200 import ops, _lib
from amq_b200._lib import PRO_RMSNORM
dev = torch.device("cuda"); H = 4096
bits = [int(b) for b in os.environ.get("BITS", "2,3,4").split(",")]
N = int(os.environ.get("N", "4096"))
keep = []
def group():
    ps = []
    for j, b in enumerate(bits):
        w = torch.randint(0, 256, (ops.native_bytes(b, N, H),), dtype=torch.uint8, device=dev); keep.append(w)
        p = ops.make_problem(b, w, h, out, N, H, prologue=PRO_RMSNORM, gamma=gamma, eps=1e-5, ldy=len(bits) * N)
        p.y = out.data_ptr() + 2 * j * N
        ps.append(p)
    return ps
h = torch.randn(1, H, device=dev).half(); gamma = torch.ones(H, device=dev).half(); out = torch.zeros(1, len(bits) * N, device=dev).half()
ws = ops.workspace(dev); L = _lib.lib()
gs = [group() for _ in range(10)]
for g in gs[:6]:
    ops.gemv_grouped(g, ws, pdl=True)
torch.cuda.synchronize()
bufs = []
for g in gs[6:]:
    b = torch.zeros(148 * 16, dtype=torch.int64, device=dev); bufs.append(b)
    L.amqb_debug_set_timeline(ctypes.c_void_p(b.data_ptr()))
    ops.gemv_grouped(g, ws, pdl=True)
L.amqb_debug_set_timeline(None)
torch.cuda.synchronize()
d = bufs[-1].cpu().view(148, 16).double(); d = d[d[:, 0] > 0]
d = torch.where(d > 0, (d - d[:, 0:1]) / 1.965e3 + 1e-6, torch.zeros_like(d))      # us since the CTA's own entry (SM clock, 1965 MHz)
names = {0: "entry", 1: "pdl_wait done", 2: "exit"}
for p in range(len(bits)):
    names.update({4 + 4 * p: f"p{p} start", 5 + 4 * p: f"p{p} x' built", 6 + 4 * p: f"p{p} records done", 7 + 4 * p: f"p{p} deposited"})
for k in sorted(names, key=lambda k: (k if k != 2 else 99)):
    c = d[:, k]
    c = c[c > 0]
    if len(c): print(f"  {names[k]:20s} min {c.min():6.2f} med {c.median():6.2f} max {c.max():6.2f}  (n={len(c)})")
