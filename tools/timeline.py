"""Per-CTA timeline of one decode-GEMV launch (globaltimer stamps), for finding where a launch's time goes."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200 import ops, _lib  # noqa: E402

bits = int(os.environ.get("BITS", "3")); N = int(os.environ.get("N", "4096")); K = int(os.environ.get("K", "4096")); M = int(os.environ.get("M", "1"))
dev = torch.device("cuda")
nb = ops.native_bytes(bits, N, K)
pool = [torch.randint(0, 256, (nb,), dtype=torch.uint8, device=dev) for _ in range(40)]
x = torch.randn(M, K, device=dev).half(); y = torch.empty(M, N, device=dev, dtype=torch.float16)
ws = ops.workspace(dev, N, K, M)
for w in pool[:20]:
    ops.gemv_grouped([ops.make_problem(bits, w, x, y, N, K)], ws)
torch.cuda.synchronize()
L = _lib.lib()
print(f"--- bits {bits} N {N} K {K} M {M}")
for trial in range(2):
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    L.amqb_debug_set_timeline(ctypes.c_void_p(dbg.data_ptr()))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.gemv_grouped([ops.make_problem(bits, pool[20 + trial], x, y, N, K)], ws)
    e1.record()
    torch.cuda.synchronize()
    L.amqb_debug_set_timeline(None)
    d = dbg.cpu().view(148, 16).double()
    d = d[d[:, 0] > 0]
    t0 = d[:, 0:1]                     # SM clocks are per-SM: everything relative to the CTA's own entry stamp
    d = torch.where(d > 0, (d - t0) / 1.965 + 1e-3, torch.zeros_like(d))   # ns at 1965 MHz
    t0 = 0.0
    names = {0: "entry", 1: "after pdl_wait", 4: "problem start", 3: "builder entry", 14: "x data arrived", 5: "x' built", 8: "round 0 done", 9: "round 1 done", 10: "round 2 done",
             11: "round 3 done", 12: "round 4 done", 13: "round 5 done", 6: "last block: records done", 7: "last block: deposited", 2: "consumer exit"}
    print(f"trial {trial}: event time {e0.elapsed_time(e1)*1e3:.1f} us, {d.shape[0]} CTAs")
    for i, nme in names.items():
        col = d[:, i]
        col = col[col > 0] - t0
        if col.numel() == 0:
            continue
        print(f"  {nme:26s} n={col.numel():3d} min {col.min()/1e3:7.2f}  median {col.median()/1e3:7.2f}  max {col.max()/1e3:7.2f} us")
