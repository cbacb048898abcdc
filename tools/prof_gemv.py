"""Tiny driver for ncu: a handful of decode-GEMV launches over distinct weight buffers."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200 import ops  # noqa: E402

bits = int(os.environ.get("BITS", "3"))
N = int(os.environ.get("N", "4096"))
K = int(os.environ.get("K", "4096"))
M = int(os.environ.get("M", "1"))
dev = torch.device("cuda")
nb = ops.native_bytes(bits, N, K)
pool = [torch.randint(0, 256, (nb,), dtype=torch.uint8, device=dev) for _ in range(24)]
x = torch.randn(M, K, device=dev).half()
y = torch.empty(M, N, device=dev, dtype=torch.float16)
ws = ops.workspace(dev, N, K, M)
for w in pool:
    ops.gemv_grouped([ops.make_problem(bits, w, x, y, N, K)], ws, pdl=False)
torch.cuda.synchronize()
