"""Back-to-back decode-GEMV launches (PDL on) with per-CTA globaltimer stamps: shows how consecutive
kernels overlap and where each one's time goes."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200 import ops, _lib  # noqa: E402

bits = int(os.environ.get("BITS", "3")); N = int(os.environ.get("N", "11008")); K = int(os.environ.get("K", "4096")); M = 1
NL = 6
dev = torch.device("cuda")
nb = ops.native_bytes(bits, N, K)
pool = [torch.randint(0, 256, (nb,), dtype=torch.uint8, device=dev) for _ in range(30)]
x = torch.randn(M, K, device=dev).half(); y = torch.empty(M, N, device=dev, dtype=torch.float16)
ws = ops.workspace(dev)
L = _lib.lib()
for w in pool[:12]:
    ops.gemv_grouped([ops.make_problem(bits, w, x, y, N, K)], ws, pdl=True)
torch.cuda.synchronize()
bufs = [torch.zeros(148 * 8, dtype=torch.int64, device=dev) for _ in range(NL)]
for i in range(NL):
    L.amqb_debug_set_timeline(ctypes.c_void_p(bufs[i].data_ptr()))
    ops.gemv_grouped([ops.make_problem(bits, pool[12 + i], x, y, N, K)], ws, pdl=True)
L.amqb_debug_set_timeline(None)
torch.cuda.synchronize()
t0 = None
names = ["entry", "pdl_wait done", "x' built", "last blk records done", "last blk stored", "-", "exit"]
for i in range(NL):
    d = bufs[i].cpu().view(148, 8).double()
    d = d[d[:, 0] > 0]
    if t0 is None:
        t0 = d[:, 0].min()
    print(f"launch {i}: ctas {d.shape[0]}  warp0: wait {d[:,5].median():.0f} cyc, proc {(d[:,7]//1000).median():.0f} cyc over {(d[:,7]%1000).median():.0f} stages; warp15 proc {d[:,3].median():.0f} red {d[:,4].median():.0f}")
    for j, nme in enumerate(names):
        if j in (3, 4, 5):
            continue
        c = (d[:, j] - t0) / 1e3
        print(f"   {nme:24s} min {c.min():7.2f} med {c.median():7.2f} max {c.max():7.2f}")
