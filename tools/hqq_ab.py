import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
from amq_b200 import ops
out = sys.argv[1]
res = {}
for bits in (2, 3, 4):
    for (N, K, seed, scale) in [(1024, 4096, 0, 0.02), (3584, 3584, 1, 0.05), (512, 1024, 2, 1.0), (256, 512, 3, 1e-3)]:
        torch.manual_seed(seed)
        W = (torch.randn(N, K, device="cuda") * scale).half()
        if seed == 2: W[0, :128] = 0; W[1, :128] = 65504.0; W[2, :64] = 6e-8
        c, s, z, it, Wq = ops.hqq_quantize(W, bits, 128, solver_dtype=torch.float16, packed=True)
        res[f"{bits}_{N}_{K}_c"] = c.cpu().numpy(); res[f"{bits}_{N}_{K}_z"] = z.cpu().numpy(); res[f"{bits}_{N}_{K}_it"] = it.cpu().numpy()
np.savez(out, **res)
print("saved", out, len(res))
