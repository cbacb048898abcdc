"""Short-prompt linears through the tcgen05 GEMM with and without the K split (AMQB_TC_NO_SPLITK=1), and the prompt pass."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gemm_tc import run
for (N, K, M) in [(4096, 4096, 63), (11008, 4096, 63), (4096, 11008, 63), (1024, 4096, 63), (1024, 4096, 512), (4096, 4096, 200)]:
    for env in ("1", None):
        if env: os.environ["AMQB_TC_NO_SPLITK"] = env
        else: os.environ.pop("AMQB_TC_NO_SPLITK", None)
        print("no split " if env else "split    ", end="")
        run(N, K, M, 3, time_it=True)
