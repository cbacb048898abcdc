"""Where config 4's time goes: Quantizer.quantize (solver + pack) vs dequantize for one linear, iterations run."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amq_b200
from amq_b200 import ops
N, K = int(os.environ.get("N", 3584)), int(os.environ.get("K", 3584))
torch.manual_seed(0)
W = (torch.randn(N, K, device="cuda") * 0.02).half()
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for bits in (2, 3, 4):
    cfg = amq_b200.BaseQuantizeConfig(nbits=bits, group_size=128)["weight_quant_params"]
    W_q, meta = amq_b200.Quantizer.quantize(W, device="cuda", compute_dtype=torch.float16, **cfg)
    meta16 = dict(meta, scale=meta["scale"].half(), zero=meta["zero"].half())
    tq = t(lambda: amq_b200.Quantizer.quantize(W, device="cuda", compute_dtype=torch.float16, **cfg))
    td = t(lambda: amq_b200.Quantizer.dequantize(W_q, meta16))
    codes, scale, zero, iters = ops.hqq_quantize(W, bits) if hasattr(ops, "hqq_quantize") else (None, None, None, None)
    ts = t(lambda: ops.hqq_quantize(W, bits))
    t0 = time.perf_counter()
    for _ in range(10): amq_b200.Quantizer.quantize(W, device="cuda", compute_dtype=torch.float16, **cfg)
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 10 * 1e6
    print(f"{N}x{K} {bits}-bit: quantize {tq:.0f} us (solver op alone {ts:.0f} us, iters {iters}), dequantize {td:.0f} us, quantize wall {wall:.0f} us")
