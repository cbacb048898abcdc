"""ncu driver: grouped q|k|v launches (RMSNorm prologue, mixed bits) over distinct weights."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200 import ops
from amq_b200._lib import PRO_RMSNORM
dev = torch.device("cuda"); H = 4096
bits = [int(b) for b in os.environ.get("BITS", "3,3,3").split(",")]
N = int(os.environ.get("N", "4096"))
h = torch.randn(1, H, device=dev).half(); gamma = torch.ones(H, device=dev).half(); out = torch.zeros(1, len(bits) * N, device=dev).half()
ws = ops.workspace(dev); keep = []
for it in range(16):
    ps = []
    for j, b in enumerate(bits):
        w = torch.randint(0, 256, (ops.native_bytes(b, N, H),), dtype=torch.uint8, device=dev); keep.append(w)
        p = ops.make_problem(b, w, h, out, N, H, prologue=PRO_RMSNORM, gamma=gamma, eps=1e-5, ldy=len(bits) * N)
        p.y = out.data_ptr() + 2 * j * N
        ps.append(p)
    ops.gemv_grouped(ps, ws, pdl=False)
torch.cuda.synchronize()
