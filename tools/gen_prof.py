import sys, os, time, torch
sys.path.insert(0, os.getcwd())
from amq_b200.arch import MODELS, sample_arch
from amq_b200.model import QuantDecoder
shape = MODELS["Llama-2-7b-hf"]; arch = sample_arch(shape, 3.0, seed=0)
m = QuantDecoder(shape, arch, batch=1, max_seq=768, seed=0)
ids = torch.randint(0, shape.vocab - 1, (1, 64))
for _ in range(3): m.generate(ids, 8)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T(); m.generate(ids, 128); t1 = T()
    # pieces
    a = T(); m.reset(); d = ids.to(m.dev); b = T()
    m.prefill(d[:, :63]); c = T()
    m.tokens.copy_(d[:, 63]); m.log_pos.zero_()
    for _ in range(128): m.step()
    e = T()
    print(f"generate {1e3*(t1-t0):.2f} ms | reset+h2d {1e3*(b-a):.3f} | prefill {1e3*(c-b):.3f} | 128 steps {1e3*(e-c):.3f} ({1e3*(e-c)/128:.4f}/step)")
# step loop without sync between: host time per replay
t0 = time.perf_counter()
for _ in range(128): m.step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host issue of 128 replays {1e3*(t1-t0):.2f} ms, + drain {1e3*(t2-t1):.2f} ms")
