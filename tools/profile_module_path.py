"""cProfile of the module path's host side (one GPTQLinear.forward call at M = 1, repeated)."""
import cProfile, pstats, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amq_b200
dev = "cuda"
K = N = 4096
m = amq_b200.GPTQLinear(3, 128, K, N, bias=False).to(dev)
m.qweight = torch.randint(-2 ** 31, 2 ** 31 - 1, m.qweight.shape, dtype=torch.int32, device=dev)
m.scales = torch.full_like(m.scales, 0.015); m.zeros = torch.full_like(m.zeros, 0.05)
x = torch.randn(1, 1, K, device=dev).half()
with torch.inference_mode():
    for _ in range(10): m(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000): m(x)
    torch.cuda.synchronize()
    print("us per forward:", (time.perf_counter() - t0) / 2000 * 1e6)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(2000): m(x)
    pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
