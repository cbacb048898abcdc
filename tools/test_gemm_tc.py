"""Quick correctness + timing probe of the tcgen05 prefill kernel against a torch fp32 reference."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200 import ops

def run(N, K, M, bits, seed=0, time_it=False):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(seed)
    codes = torch.randint(0, 2 ** bits, (N, K), device=dev, dtype=torch.uint8, generator=g)
    lo, hi = {2: (0.024, 0.054), 3: (0.010, 0.023), 4: (0.0047, 0.011)}[bits]
    scale = torch.empty(N, K // 128, device=dev).uniform_(lo, hi, generator=g).half()
    zero = torch.empty(N, K // 128, device=dev).uniform_(0.5, 2 ** bits - 1.5, generator=g).half()
    nat = ops.pack_native(bits, codes, scale, zero)
    W = (codes.float().reshape(N, K // 128, 128) * scale.float()[..., None] - (zero * scale).float()[..., None]).reshape(N, K)
    x = torch.randn(M, K, device=dev, generator=g).half()
    bias = torch.randn(N, device=dev, generator=g).half()
    ref = x.float() @ W.t() + bias.float()
    y = ops.gemm_tc(bits, nat, x, N, K, bias)
    torch.cuda.synchronize()
    rel = float((y.float() - ref).abs().max() / ref.abs().max())
    msg = f"N={N} K={K} M={M} bits={bits} max-rel {rel:.2e}"
    if time_it:
        # device time: 20 calls captured in one CUDA graph (a python loop of ~30 us calls would be host-bound)
        ws = ops.gemm_workspace(M, K, bits, x.device) if hasattr(ops, "gemm_workspace") else None
        y2 = torch.empty(M, N, device=dev, dtype=torch.float16)
        def call():
            ops.gemm_tc(bits, nat, x, N, K, bias, out=y2, workspace=ws)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                call()
            s.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for _ in range(20):
                    call()
            gr.replay(); s.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(5):
                gr.replay()
            e1.record(s); s.synchronize()
        ms = e0.elapsed_time(e1) / 100
        msg += f"  {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s"
    print(msg, flush=True)
    return rel

if __name__ == "__main__":
    bad = 0
    for bits in (4, 2, 3):
        for (N, K, M) in [(128, 128, 128), (128, 256, 64), (256, 512, 40), (512, 1024, 300)]:
            bad += run(N, K, M, bits) > 1e-3
    if len(sys.argv) > 1:
        for bits in (2, 3, 4):
            run(4096, 4096, 512, bits, time_it=True)
            run(11008, 4096, 512, bits, time_it=True)
            run(4096, 4096, 2048, bits, time_it=True)
    print("FAILED" if bad else "ALL OK")
