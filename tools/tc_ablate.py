"""Ablation timing of the tcgen05 prefill kernel (AMQB_TC_DBG mask); results are wrong by construction."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200 import ops
from amq_b200.model import synthetic_native
dev = torch.device("cuda"); bits = 3; N = K = 4096
g = torch.Generator(device=dev).manual_seed(0)
w = synthetic_native(bits, N, K, dev, g)
for M in (512, 2048):
    x = torch.randn(M, K, device=dev).half()
    ws = ops.gemm_workspace(M, K, bits, dev); y = torch.empty(M, N, device=dev, dtype=torch.float16)
    for mask in (0, 1, 2, 4, 3, 5, 6, 7):
        os.environ["AMQB_TC_DBG"] = str(mask)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                ops.gemm_tc(bits, w, x, N, K, out=y, workspace=ws)
            s.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for _ in range(20):
                    ops.gemm_tc(bits, w, x, N, K, out=y, workspace=ws)
            gr.replay(); s.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(5):
                gr.replay()
            e1.record(s); s.synchronize()
        print(f"M={M} dbg={mask} (1:no-deq 2:no-mma 4:no-x)  {e0.elapsed_time(e1) * 10:.1f} us", flush=True)
