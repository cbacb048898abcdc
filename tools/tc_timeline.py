import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200 import ops
from amq_b200.model import synthetic_native
dev = torch.device("cuda"); bits = 3; N = K = 4096; M = 512
g = torch.Generator(device=dev).manual_seed(0)
w = synthetic_native(bits, N, K, dev, g)
x = torch.randn(M, K, device=dev).half()
for mask in [int(a) for a in sys.argv[1:]] or (8, 8, 8 + 7, 8 + 1, 8 + 2):
    os.environ["AMQB_TC_DBG"] = str(mask)
    print("dbg", mask, flush=True)
    ops.gemm_tc(bits, w, x, N, K); torch.cuda.synchronize()
