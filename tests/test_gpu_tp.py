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
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


def test_allreduce_and_row_parallel(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29621", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == n
