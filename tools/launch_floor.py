"""Per-launch floor of dependent kernel chains inside a CUDA graph (with / without PDL): a trivial glue
kernel, and the decode GEMV on a tiny problem (fixed cost of the kernel itself)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200 import ops, _lib
from amq_b200._lib import check, cur_stream, lib, ptr

dev = torch.device("cuda")
L = lib()
tok = torch.zeros(1, dtype=torch.int64, device=dev)
table = torch.zeros(16, 4096, dtype=torch.float16, device=dev)
out = torch.zeros(1, 4096, dtype=torch.float16, device=dev)

def chain(fn, n, pdl):
    L.amqb_set_pdl(int(pdl))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn(pdl)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn(pdl)
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10):
            g.replay()
        e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (10 * n)

def embed(pdl):
    check(L.amqb_embed(ptr(tok), ptr(table), ptr(out), 1, 4096, cur_stream()))

for pdl in (False, True):
    print(f"trivial kernel chain, pdl={pdl}: {chain(embed, 200, pdl):.2f} us per launch")

ws = ops.workspace(dev)
for (N, K, bits) in [(4736, 128, 3), (4736, 512, 3), (4096, 4096, 3), (4096, 4096, 2)]:
    nat = torch.randint(0, 256, (ops.native_bytes(bits, N, K),), dtype=torch.uint8, device=dev)
    x = torch.randn(1, K, device=dev).half(); y = torch.empty(1, N, device=dev, dtype=torch.float16)
    p = ops.make_problem(bits, nat, x, y, N, K)
    for pdl in (False, True):
        t = chain(lambda q: ops.gemv_grouped([p], ws, pdl=q), 100, pdl)
        print(f"gemv N={N} K={K} bits={bits} (L2-resident weights) pdl={pdl}: {t:.2f} us per launch")
