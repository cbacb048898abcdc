import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200 import ops
from amq_b200.model import synthetic_native
dev = torch.device("cuda"); bits = int(os.environ.get("BITS", "3")); N = K = 4096; M = int(os.environ.get("M", "512"))
g = torch.Generator(device=dev).manual_seed(0)
w = synthetic_native(bits, N, K, dev, g)
x = torch.randn(M, K, device=dev).half()
for _ in range(6):
    ops.gemm_tc(bits, w, x, N, K)
torch.cuda.synchronize()
