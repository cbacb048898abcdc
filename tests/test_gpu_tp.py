"""Tensor-parallel path on >= 2 GPUs (run with `gpurun --gpus 2`): the one-shot NVLink all-reduce
against ncclAllReduce, and a row-parallel quantized linear (reference-layout buffers sliced per rank,
repacked, decode GEMV, all-reduce) against the unsharded result."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from amq_b200 import ops, tp
from oracle import amq_oracle as O
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
H = 8192
ar = tp.PeerAllReduce(rank, world, H * 4, pdl=False)
# 1. collective vs NCCL, several rounds (epoch / parity logic), with and without residual
for it in range(6):
    n = H * (1 + it % 3)
    torch.manual_seed(100 * it + rank)
    part = torch.randn(n, device=dev).half()
    torch.manual_seed(7 + it)
    h = torch.randn(n, device=dev).half()
    ref = part.float().clone()
    dist.all_reduce(ref)
    ref = (ref + h.float())
    out = h.clone()
    ar(part, out)
    torch.cuda.synchronize()
    # fp32 accumulation in rank order on every rank: one rounding to fp16
    assert torch.equal(out, ref.half()) or (out.float() - ref).abs().max() <= 2e-3 * ref.abs().max(), it
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    assert all(torch.equal(g, gathered[0]) for g in gathered)       # bit-identical on all ranks
print("collective ok-ish", rank, flush=True)
# 2. row-parallel 3-bit linear
rs = np.random.RandomState(0)
N, K, bits = 256, 1024, 3
codes = rs.randint(0, 8, size=(N, K))
qw = torch.from_numpy(O.gptq_pack_codes(codes, bits))
sc = torch.from_numpy(rs.uniform(0.01, 0.02, size=(K // 128, N)).astype(np.float32)).half().float()
ze = torch.from_numpy(rs.uniform(0.02, 0.1, size=(K // 128, N)).astype(np.float32)).half().float()
x = torch.from_numpy(rs.randn(1, K).astype(np.float32)).half()
q, s, z = tp.shard_gptq_buffers(qw, sc, ze, bits, "row", rank, world)
Kl = K // world
nat = ops.repack_gptq(bits, q.to(dev), s.to(dev), z.to(dev), N, Kl, 128)
part = ops.gemv(bits, nat, x[:, rank * Kl:(rank + 1) * Kl].contiguous().to(dev), N, Kl).reshape(-1)
pad = torch.zeros(H, device=dev, dtype=torch.float16); pad[:N] = part
out = torch.zeros(H, device=dev, dtype=torch.float16)
ar(pad, out)
torch.cuda.synchronize()
full = O.gptq_forward_fp32(x, qw.numpy(), sc, ze, bits, 128).reshape(-1)
assert O.max_rel(out[:N].cpu(), full) <= 2e-3, O.max_rel(out[:N].cpu(), full)
# 3. model level: the sharded decoder (every rank adopts its shard of the SAME full model) against the unsharded one
from amq_b200.arch import ModelShape, LINEARS
from amq_b200.model import QuantDecoder
shape = ModelShape("tiny-gqa", 512, 1024, 8, 4, 2, 512, head_dim=64)
rs = np.random.RandomState(3)
arch = {n: rs.choice([2, 3, 4], size=2).tolist() for n in LINEARS}
full = QuantDecoder(shape, arch, batch=1, max_seq=32, device=f"cuda:{local}", seed=11)
for kind in ("fused", "amqb", "nccl"):
    print("model-level", kind, rank, flush=True)
    m = QuantDecoder(shape, arch, batch=1, max_seq=32, device=f"cuda:{local}", seed=0, tp_rank=rank, tp_world=world)
    m.adopt_shard_of(full)
    if kind == "nccl":
        m.attach_allreduce(tp.NcclAllReduce(), fused=False)
    else:
        m.attach_allreduce(tp.PeerAllReduce(rank, world, shape.hidden), fused=(kind == "fused"))
    tok = torch.tensor([17], device=dev)
    for mm in (full, m):
        mm.reset(); mm.tokens.copy_(tok)
    for pos in range(6):
        print(" pos", pos, rank, flush=True)
        m.tokens.copy_(full.tokens)              # same token stream on both
        full.step(); m.step()                    # graph-replayed (captured on first call)
        torch.cuda.synchronize()
        rel = float((m.logits - full.logits).abs().max() / full.logits.abs().max())
        assert rel <= 2e-2, (kind, pos, rel)
    lg = [torch.empty_like(m.logits) for _ in range(world)]
    dist.all_gather(lg, m.logits)
    assert all(torch.equal(g, lg[0]) for g in lg), kind          # every rank holds bit-identical logits
    dist.barrier()
    if kind != "nccl":
        m.allreduce.close()                                      # collective: unmap, barrier, free
# 4. all-reduce of an activation matrix (prompt pass) against NCCL, several sizes / rounds (per-launch parity)
arr = tp.PeerAllReduce(rank, world, 8, pdl=False, rows_elems=1 << 20)
for it, n in enumerate([8, 8200, 1 << 20, 512 * 40, 1 << 20, 8]):
    torch.manual_seed(100 * it + rank)
    part = torch.randn(n, device=dev).half()
    torch.manual_seed(7 + it)
    h = torch.randn(n, device=dev).half()
    ref = part.float().clone()
    dist.all_reduce(ref)
    ref = ref + h.float()
    out = h.clone()
    arr.rows(part, out)
    torch.cuda.synchronize()
    assert (out.float() - ref).abs().max() <= 2e-3 * ref.abs().max(), ("rows", it)
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
arr.close()
# 5. the sharded decoder's prompt pass (plain launches, then the captured graph) against the unsharded one
S = 64
full = QuantDecoder(shape, arch, batch=2, max_seq=S, device=f"cuda:{local}", seed=11)
m = QuantDecoder(shape, arch, batch=2, max_seq=S, device=f"cuda:{local}", seed=0, tp_rank=rank, tp_world=world)
m.adopt_shard_of(full)
m.attach_allreduce(tp.PeerAllReduce(rank, world, shape.hidden * 2, rows_elems=shape.hidden * 2 * S), fused=True)
torch.manual_seed(5)
ids = torch.randint(0, shape.vocab, (2, 41), device=dev)
hk = full.Hkv // world
for use_graph in (False, True, True):
    for mm in (full, m):
        mm.reset()
        for L in mm.layers:
            L["k_cache"].zero_(); L["v_cache"].zero_()
    full.prefill(ids[:, :40]); m.prefill(ids[:, :40], use_graph=use_graph)
    torch.cuda.synchronize()
    for Lf, Lm in zip(full.layers, m.layers):
        for c in ("k_cache", "v_cache"):
            want = Lf[c][:, rank * hk:(rank + 1) * hk, :40].float(); got = Lm[c][:, :, :40].float()
            assert (want - got).abs().max() <= 1e-2 * want.abs().max(), (use_graph, c)
    for mm in (full, m):
        mm.tokens.copy_(ids[:, 40]); mm.step()
    torch.cuda.synchronize()
    rel = float((m.logits - full.logits).abs().max() / full.logits.abs().max())
    assert rel <= 2e-2, ("prefill", use_graph, rel)
dist.barrier()
m.allreduce.close()
dist.barrier()
torch.cuda.synchronize()
print("worker-done", rank, flush=True)
# CUDA graphs that captured NCCL kernels are still alive: tearing the communicator down under them can block forever
os._exit(0)
'''


def test_allreduce_and_row_parallel(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29621", str(script), ROOT]
    log = tmp_path / "out.txt"
    with open(log, "w") as f:
        p = subprocess.Popen(cmd, stdout=f, stderr=subprocess.STDOUT, text=True)
        try:
            rc = p.wait(timeout=200)
        except subprocess.TimeoutExpired:
            p.kill()
            rc = -9
    out = log.read_text()
    assert rc == 0, out[-4000:]
    assert out.count("worker-done") == n   # (the ranks' lines interleave in the shared log; earlier stages print "ok" too)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("batch", [1, 2])
def test_tp_decoder_matches_unsharded_emulated(world, batch, fused):
    """Model-level tensor-parallel parity on ONE GPU: `world` emulated ranks (tp.LocalTPGroup: one stream
    per rank, each with its Megatron shard of the same weights and its own heads' K/V cache, the one-shot all-reduce
    meeting through device memory) against the unsharded decoder: logits within 2e-2 (fp16 activations; the shards
    round their partial sums separately), bit-identical on every rank, eager and graph-replayed."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import numpy as np
    from amq_b200 import tp
    from amq_b200.arch import LINEARS, ModelShape
    from amq_b200.model import QuantDecoder
    shape = ModelShape("tiny-gqa", 512, 1024, 8, 4, 2, 512, head_dim=64, qkv_bias=(batch == 2))
    rs = np.random.RandomState(world)
    arch = {n: rs.choice([2, 3, 4], size=2).tolist() for n in LINEARS}
    full = QuantDecoder(shape, arch, batch=batch, max_seq=32, seed=5)
    grp = tp.LocalTPGroup(full, world, fused=fused)      # fused: all-reduce inside the row-parallel GEMVs' epilogue
    tok = torch.randint(0, shape.vocab, (batch,), device=full.dev)
    full.reset(); full.tokens.copy_(tok)
    grp.set_tokens(tok)
    for pos in range(8):
        for m in grp.ranks:
            m.tokens.copy_(full.tokens)
        torch.cuda.synchronize()
        if pos < 2:
            full.step_eager(); grp.step_eager()
        else:
            full.step(); grp.step()
        torch.cuda.synchronize()
        ref = full.logits
        for m in grp.ranks:
            assert torch.equal(m.logits, grp.ranks[0].logits)
        rel = float((grp.ranks[0].logits - ref).abs().max() / ref.abs().max())
        assert rel <= 2e-2, (world, batch, pos, rel)
    assert grp.timeouts() == 0


def _emulated_ranks_allreduce_rows(world, sizes):
    import ctypes
    from amq_b200 import tp
    from amq_b200._lib import lib
    dev = torch.device("cuda", 0)
    max_elems = max(sizes)
    nbytes = int(lib().amqb_ar_rows_buffer_bytes(max_elems, world))
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    ars = [tp.LocalAllReduce(r, world, 8, bufs, pdl=False, rows_elems=max_elems, rows_bufs=bufs) for r in range(world)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    g = torch.Generator(device=dev).manual_seed(world)
    for it, n in enumerate(sizes):
        parts = [torch.randn(n, device=dev, generator=g).half() for _ in range(world)]
        h = torch.randn(n, device=dev, generator=g).half()
        outs = [h.clone() for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):                       # every rank's launch is issued before anything is waited for
            with torch.cuda.stream(streams[r]):
                ars[r].rows(parts[r], outs[r])
        torch.cuda.synchronize()
        acc = torch.zeros(n, device=dev)
        for r in range(world):                       # rank order, fp32, residual last: what the kernel does
            acc = acc + parts[r].float()
        want = (acc + h.float()).half()
        for r in range(world):
            assert torch.equal(outs[r], want), (world, it, n, r)
    for b in bufs:
        c = ctypes.c_int(0)
        assert lib().amqb_ar_rows_timeouts(ctypes.c_void_p(b.data_ptr()), ctypes.byref(c)) == 0 and c.value == 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_allreduce_rows_emulated(world):
    """amqb_allreduce_rows_f16 (the prompt pass's [B*T, hidden] all-reduce: one-shot push spread over up to 64 CTAs, slot
    parity per launch) with `world` emulated ranks on one device, one stream each: exact fp32 rank-order sum, identical on
    every rank, over sizes that use one CTA, a ragged last CTA, all 64 CTAs, and back (stale slots of a larger launch)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    # (emulated ranks share one device and a waiting CTA keeps its slot: all ranks' CTAs must fit the device at once)
    big = (1 << 20) if world <= 4 else (1 << 18)
    _emulated_ranks_allreduce_rows(world, [8, 8200, big, 512 * 40, big - 8, 8, 1024 * 8 * 3 + 16])


def test_allreduce_rows_rejects_bad_arguments():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ctypes
    from amq_b200._lib import cur_stream, lib, ptr
    L = lib()
    assert L.amqb_ar_rows_buffer_bytes(0, 2) == 0 and L.amqb_ar_rows_buffer_bytes(8, 17) == 0
    buf = torch.zeros(int(L.amqb_ar_rows_buffer_bytes(64, 1)), dtype=torch.uint8, device="cuda")
    peers = (ctypes.c_void_p * 1)(ctypes.c_void_p(buf.data_ptr()))
    x = torch.zeros(64, dtype=torch.float16, device="cuda")
    call = lambda n, mx: L.amqb_allreduce_rows_f16(peers, 0, 1, ptr(x), ptr(x), ptr(x), ctypes.c_longlong(n),
                                                    ctypes.c_longlong(mx), 0, cur_stream())
    assert call(12, 64) != 0 and call(128, 64) != 0 and call(0, 64) != 0
    assert call(64, 64) == 0                         # world = 1: out = residual + partial
    torch.cuda.synchronize()


@pytest.mark.parametrize("world,batch,prompt", [(2, 2, 6), (2, 1, 41), (4, 2, 30), (8, 1, 41)])
def test_tp_prefill_matches_unsharded_emulated(world, batch, prompt):
    """The tensor-parallel prompt pass (QuantDecoder.prefill with tp_world > 1: column-parallel q|k|v / gate|up on this
    rank's heads and columns, row-parallel o_proj / down_proj followed by the [B*T, hidden] all-reduce) with emulated
    ranks against the unsharded prompt pass: every rank's K/V cache equals its heads of the unsharded cache (1e-2), the
    next decode step's logits agree (2e-2) and are bit-identical across ranks.  10 rows take the skinny decode kernel,
    40 / 58 rows the tcgen05 GEMM."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import numpy as np
    from amq_b200 import tp
    from amq_b200.arch import LINEARS, ModelShape
    from amq_b200.model import QuantDecoder
    if world == 8:          # 16 / 8 heads: two query heads and one K/V head per rank, 128 of the MLP's columns
        shape = ModelShape("tiny-gqa8", 1024, 1024, 16, 8, 2, 512, head_dim=64)
    else:
        shape = ModelShape("tiny-gqa", 512, 1024, 8, 4, 2, 512, head_dim=64, qkv_bias=(batch == 2))
    rs = np.random.RandomState(world)
    arch = {n: rs.choice([2, 3, 4], size=2).tolist() for n in LINEARS}
    full = QuantDecoder(shape, arch, batch=batch, max_seq=64, seed=5)
    grp = tp.LocalTPGroup(full, world, fused=True)
    ids = torch.randint(0, shape.vocab, (batch, prompt), device=full.dev)
    P = prompt - 1
    full.reset()
    grp.set_tokens(ids[:, 0])
    full.prefill(ids[:, :P])
    grp.prefill(ids[:, :P])
    torch.cuda.synchronize()
    hk = full.Hkv // world
    for r, m in enumerate(grp.ranks):
        assert int(m.pos.item()) == P and m._pos_h == P
        for Lf, Lm in zip(full.layers, m.layers):
            for c in ("k_cache", "v_cache"):
                want, got = Lf[c][:, r * hk:(r + 1) * hk, :P].float(), Lm[c][:, :, :P].float()
                assert (want - got).abs().max() <= 1e-2 * want.abs().max(), (world, r, c)
    full.tokens.copy_(ids[:, P])
    for m in grp.ranks:
        m.tokens.copy_(ids[:, P])
    torch.cuda.synchronize()
    full.step_eager(); grp.step_eager()
    torch.cuda.synchronize()
    for m in grp.ranks:
        assert torch.equal(m.logits, grp.ranks[0].logits)
    rel = float((grp.ranks[0].logits - full.logits).abs().max() / full.logits.abs().max())
    assert rel <= 2e-2, (world, batch, prompt, rel)
    assert grp.timeouts() == 0


def test_tp_decoder_split_attention_emulated(monkeypatch):
    """Tensor-parallel decoders take the split-KV decode attention like the unsharded one (each rank over its own heads):
    forced on from position 8, eager and graph-replayed across the threshold, against the unsharded decoder."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import numpy as np
    from amq_b200 import tp
    from amq_b200.arch import LINEARS, ModelShape
    from amq_b200.model import QuantDecoder
    monkeypatch.setenv("AMQB_ATTN_SPLIT_MIN_POS", "8")
    shape = ModelShape("tiny-gqa", 512, 1024, 8, 4, 2, 512, head_dim=64)
    rs = np.random.RandomState(9)
    arch = {n: rs.choice([2, 3, 4], size=2).tolist() for n in LINEARS}
    full = QuantDecoder(shape, arch, batch=1, max_seq=32, seed=5)
    grp = tp.LocalTPGroup(full, 2, fused=True)
    assert all(m.attn_splits == 4 and m.attn_split_min_pos == 8 for m in grp.ranks)
    tok = torch.randint(0, shape.vocab, (1,), device=full.dev)
    full.reset(); full.tokens.copy_(tok)
    grp.set_tokens(tok)
    for pos in range(16):
        for m in grp.ranks:
            m.tokens.copy_(full.tokens)
        torch.cuda.synchronize()
        if pos in (0, 1, 9):
            full.step_eager(); grp.step_eager()
        else:
            full.step(); grp.step()
        torch.cuda.synchronize()
        assert all(m.graph_long is not None for m in grp.ranks) or pos < 2
        rel = float((grp.ranks[0].logits - full.logits).abs().max() / full.logits.abs().max())
        assert rel <= 2e-2, (pos, rel)
        assert torch.equal(grp.ranks[0].logits, grp.ranks[1].logits)
    assert grp.timeouts() == 0
