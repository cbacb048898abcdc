"""Where the per-token host overhead of the synchronous (host-owned) loop goes."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200.arch import MODELS, sample_arch
from amq_b200.model import QuantDecoder
shape = MODELS["Llama-2-7b-hf"]
arch = sample_arch(shape, 3.0, seed=0)
m = QuantDecoder(shape, arch, batch=1, max_seq=512, seed=0)
m.capture()
hi, ho = torch.ones(1, dtype=torch.int64).pin_memory(), torch.zeros(1, dtype=torch.int64).pin_memory()
m.step_host(hi, ho)
def wall(fn, n=200):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
st = torch.cuda.current_stream()
m.reset(); print("back-to-back replay            %.1f us/step" % wall(lambda: m.step()))
m.reset(); print("replay + stream sync           %.1f us/step" % wall(lambda: (m.step(), st.synchronize())))
m.reset(); print("step_host (io graph + sync)    %.1f us/step" % wall(lambda: m.step_host(hi, ho)))
m.reset(); print("step_host + host feedback      %.1f us/step" % wall(lambda: (m.step_host(hi, ho), hi.copy_(ho))))
ev = torch.cuda.Event(enable_timing=False)
def ev_wait():
    m.step(); ev.record(); 
    while not ev.query(): pass
m.reset(); print("replay + event spin            %.1f us/step" % wall(ev_wait))
g = torch.cuda.CUDAGraph()
x = torch.zeros(1, device="cuda")
with torch.cuda.graph(g): x.add_(1)
print("1-node graph replay + sync     %.1f us" % wall(lambda: (g.replay(), st.synchronize())))
