"""One tcgen05 GEMM call at a short prompt (M = 63): in-kernel timeline (AMQB_TC_DBG=8) next to the graph-timed call."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gemm_tc import run
os.environ["AMQB_TC_DBG"] = "8"
for (N, K, M) in [(4096, 4096, 63), (4096, 11008, 63), (4096, 4096, 511)]:
    run(N, K, M, 3)
    torch.cuda.synchronize()
os.environ.pop("AMQB_TC_DBG")
for (N, K, M) in [(4096, 4096, 63), (4096, 4096, 511)]:
    run(N, K, M, 3, time_it=True)
