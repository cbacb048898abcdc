"""Absolute (globaltimer) timeline of the decode GEMV launches of consecutive layers, as the model issues them (eager
launches, PDL on): where the time between two launches goes.  Needs an AMQB_TIMELINE=1 build of the library.
Per launch: min / median / max over the CTAs of each stamp, in us relative to the first instrumented launch."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amq_b200.arch import MODELS, sample_arch
from amq_b200.model import QuantDecoder
from amq_b200._lib import lib

shape = MODELS[os.environ.get("MODEL", "Llama-2-7b-hf")]
arch = sample_arch(shape, float(os.environ.get("BITS", "3.0")), seed=0)
nb = int(os.environ.get("BLOCKS", "6"))
m = QuantDecoder(shape, arch, batch=1, max_seq=256, n_block=nb)
m.pos.fill_(100)
L = lib()
L.amqb_set_pdl(1)
bufs = []
orig = m._gemv
names = ["qkv", "o", "gu", "down"]
# buffers are allocated BEFORE the capture: a torch.zeros inside it would put a fill kernel between the decode launches
# and break the programmatic-dependent-launch chain this tool is meant to observe
pool = [torch.zeros(148 * 16 + 148 * 32, dtype=torch.int64, device=m.dev) for _ in range(4 * nb)]
def hooked(arr, n):
    b = pool[len(bufs)]
    bufs.append((names[len(bufs) % 4], n, b))
    L.amqb_debug_set_timeline(ctypes.c_void_p(b.data_ptr()))
    orig(arr, n)
for _ in range(3):
    m._step_launches()
torch.cuda.synchronize()
m._gemv = hooked
if os.environ.get("EAGER") == "1":
    m._step_launches()
else:
    # the buffer pointer is baked into each captured launch: replaying the graph gives the back-to-back (PDL) behaviour
    st_ = torch.cuda.Stream()
    with torch.cuda.stream(st_):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st_):
            m._step_launches()
        for _ in range(3):
            g.replay()
    st_.synchronize()
torch.cuda.synchronize()
L.amqb_debug_set_timeline(None)
t0 = None
for i, (name, n, b) in enumerate(bufs):
    full_b = b.cpu()
    d = full_b[:148 * 16].view(148, 16).double()
    rbs = full_b[148 * 16:].view(148, 32).double()
    live = d[:, 0] > 0
    rbs = rbs[live]
    d = d[d[:, 0] > 0]
    gt = full_b[148 * 16:].view(148, 32)[live][:, 30]       # int64 globaltimer at entry (slot 12 is reused by the third problem's stamps)
    if t0 is None:
        t0 = int(gt.min())
    if i < 4 * (nb - 3):         # print the last three layers
        continue
    ent = (gt - t0).double() / 1e3          # us, globaltimer at entry (subtracted as integers: 1.8e18 ns does not fit a double)
    def st(col):                            # SM clocks since the CTA's own entry -> us
        c = d[:, col]; ok = c > 0
        c = (c[ok] - d[ok, 0]) / 1965.0
        return "      -" if len(c) == 0 else f"{c.min():5.2f} {c.median():5.2f} {c.max():5.2f}"
    print(f"{name:5s} ctas {d.shape[0]:3d} | entry(abs) {ent.min():7.2f} {ent.median():7.2f} {ent.max():7.2f} | since entry: init {st(13)} | pre-wait {st(1)} | "
          f"waited {st(15)} | x arrived {st(14)} | x' built {st(5)} | it0 data {st(8)} done {st(9)} | it1 data {st(10)} done {st(11)} | recs done(last p) {st(6 + 4 * (n - 1))} | cons exit {st(2)} | reducer done {st(3)}")

    smid = rbs[:, 31].long() - 1
    cnt = torch.bincount(smid[smid >= 0], minlength=148)
    endt = ent + (d[:, 3] - d[:, 0]).clamp(min=0) / 1965.0
    print(f"      SMs with 0/1/2+ CTAs of this launch: {(cnt == 0).sum().item()}/{(cnt == 1).sum().item()}/{(cnt >= 2).sum().item()} | reducer done(abs) {endt.min():7.2f} {endt.median():7.2f} {endt.max():7.2f}")
    if os.environ.get("PRODUCER") == "1":
        ent0 = d[:, 0]
        iss = rbs[:, 24:30]
        print("      producer: issue of stage 0..2, then landing of fill 0..2 (us since entry, median): " + " ".join(
            f"{((iss[:, k][iss[:, k] > 0] - ent0[iss[:, k] > 0]) / 1965.0).median():5.2f}" if (iss[:, k] > 0).any() else "    -" for k in range(6)))
    if os.environ.get("ROWBLOCKS") == "1" and name in ("gu", "qkv"):
        ent0 = d[:, 0]
        line = []
        for k in range(8):
            cols = rbs[:, 3 * k: 3 * k + 3]
            ok = cols[:, 0] > 0
            if ok.sum() == 0:
                break
            c = (cols[ok] - ent0[ok, None]) / 1965.0
            line.append(f"rb{k}[n={int(ok.sum())}] start {c[:,0].median():5.2f} recs {c[:,1].median():5.2f} dep {c[:,2].median():5.2f}")
        print("      " + " | ".join(line))
