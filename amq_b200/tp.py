"""Tensor parallelism for the only config that shards (config 5, Llama-2-70B): Megatron-style
column-parallel q/k/v/gate/up, row-parallel o/down, one all-reduce of the [M, hidden] partial sums
after each row-parallel linear.  One process per GPU; torch.distributed (NCCL) is the
rendezvous / baseline collective, the data path is amqb_allreduce_f16: a one-shot push all-reduce
over NVLink peer memory (amq_b200/csrc/allreduce.cu)."""
from __future__ import annotations

import ctypes
import json
import os
from typing import Dict, List, Optional

import torch

from .arch import MODELS, ModelShape, sample_arch

GROUP = 128


def shard_plan(shape: ModelShape, world: int) -> Dict[str, Dict[str, int]]:
    """Per-rank (N, K) of every linear and how it is split.  Column-parallel = split N,
    row-parallel = split K; group (128) and 32-row record boundaries must be respected."""
    if shape.n_heads % world or shape.n_kv_heads % world or shape.inter % (world * GROUP):
        raise ValueError(f"{shape.name} does not shard {world}-way (heads {shape.n_heads}/{shape.n_kv_heads}, inter {shape.inter})")
    q = shape.n_heads * shape.head_dim // world
    kv = shape.n_kv_heads * shape.head_dim // world
    i = shape.inter // world
    plan = {
        "self_attn.q_proj": {"N": q, "K": shape.hidden, "split": "column"},
        "self_attn.k_proj": {"N": kv, "K": shape.hidden, "split": "column"},
        "self_attn.v_proj": {"N": kv, "K": shape.hidden, "split": "column"},
        "self_attn.o_proj": {"N": shape.hidden, "K": q, "split": "row"},
        "mlp.gate_proj": {"N": i, "K": shape.hidden, "split": "column"},
        "mlp.up_proj": {"N": i, "K": shape.hidden, "split": "column"},
        "mlp.down_proj": {"N": shape.hidden, "K": i, "split": "row"},
    }
    for name, p in plan.items():
        if p["N"] % 32 or p["K"] % GROUP:
            raise ValueError(f"{name}: shard {p['N']}x{p['K']} breaks the 32-row / 128-k record grid")
    return plan


def shard_gptq_buffers(qweight: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor, bits: int, split: str,
                       rank: int, world: int, group: int = GROUP):
    """Slice reference-layout GPTQLinear buffers for one rank: column split = [:, n0:n1];
    row split = qweight rows [k0*bits/32, k1*bits/32) and scales/zeros rows [k0/G, k1/G)."""
    K = qweight.shape[0] * 32 // bits
    N = qweight.shape[1]
    if split == "column":
        n0, n1 = N * rank // world, N * (rank + 1) // world
        return qweight[:, n0:n1].contiguous(), scales[:, n0:n1].contiguous(), zeros[:, n0:n1].contiguous()
    k0, k1 = K * rank // world, K * (rank + 1) // world
    assert k0 % group == 0 and (k0 * bits) % 32 == 0
    return (qweight[k0 * bits // 32: k1 * bits // 32].contiguous(), scales[k0 // group: k1 // group].contiguous(),
            zeros[k0 // group: k1 // group].contiguous())


def exchange_handles(handle: bytes, group=None) -> List[bytes]:
    """All-gather one opaque 64-byte handle per rank over the host channel of torch.distributed."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out: List[Optional[bytes]] = [None] * world
    dist.all_gather_object(out, handle, group=group)
    return [bytes(h) for h in out]


class PeerAllReduce:
    """Owns this rank's exchange buffer and the mapped peer buffers; __call__(partial, h) does
    h <- h + sum_ranks(partial)."""

    def __init__(self, rank: int, world: int, max_elems: int, pdl: bool = True):
        from ._lib import check, lib
        self.rank, self.world, self.max_elems, self.pdl = rank, world, max_elems, pdl
        L = lib()
        nbytes = int(L.amqb_ar_buffer_bytes(max_elems, world))
        mine = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        check(L.amqb_ar_alloc(ctypes.c_size_t(nbytes), ctypes.byref(mine), handle), "ar_alloc")
        self._mine = mine
        handles = exchange_handles(bytes(handle))
        self._peers = (ctypes.c_void_p * world)()
        self._opened = []
        for r, h in enumerate(handles):
            if r == rank:
                self._peers[r] = mine
            else:
                p = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                check(L.amqb_ar_open(buf, ctypes.byref(p)), "ar_open")
                self._peers[r] = p
                self._opened.append(p)
        import torch.distributed as dist
        dist.barrier()

    def __call__(self, partial: torch.Tensor, h: torch.Tensor) -> None:
        from ._lib import check, cur_stream, lib, ptr
        check(lib().amqb_allreduce_f16(self._peers, self.rank, self.world, ptr(partial), ptr(h), ptr(h), partial.numel(),
                                       self.max_elems, int(self.pdl), cur_stream()), "allreduce")


class NcclAllReduce:
    """Baseline: ncclAllReduce on the compute stream + residual add."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist

    def __call__(self, partial: torch.Tensor, h: torch.Tensor) -> None:
        self.dist.all_reduce(partial)
        h.add_(partial)


def bench_main(args) -> None:
    """bench.py --workload llama70b-tp: Llama-2-70B random-init, AMQ avg 3.0 bits, batch-1 decode,
    tensor-parallel over --gpus ranks (strong scaling: same model, 1/tp of the weights per GPU)."""
    import torch.distributed as dist
    from .model import QuantDecoder
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shape = MODELS["Llama-2-70b-hf"]
    arch = sample_arch(shape, 3.0, seed=0)
    shard_plan(shape, world)
    n_block = int(os.environ.get("AMQB_BLOCKS", shape.n_block))
    model = QuantDecoder(shape, arch, batch=1, max_seq=max(256, args.warmup + args.steps + 8), device=f"cuda:{local}",
                         seed=0, n_block=n_block, tp_rank=rank, tp_world=world)
    ar_kind = os.environ.get("AMQB_AR", "amqb")
    if world > 1:
        model.allreduce = PeerAllReduce(rank, world, shape.hidden) if ar_kind == "amqb" else NcclAllReduce()
    model.capture()
    lps = model.launches_per_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    model.reset()
    model.tokens.fill_(1)
    for _ in range(args.warmup):
        model.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        model.step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if rank == 0:
        by = model.algorithmic_bytes_per_token()
        peak = 6549.4
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
        if os.path.exists(p):
            peak = float(json.load(open(p))["hbm_gbs"])
        tok_s = args.steps / (ms * 1e-3)
        print(json.dumps({
            "metric": "batch-1 decode tok/s, Llama-2 70B AMQ 3-bit avg, tensor parallel", "value": tok_s, "unit": "tok/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16 (fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": f"{shape.name} random-init ({n_block} blocks), AMQ avg 3.0 bits, batch-1 decode, tp={world}",
                       "allreduce": ar_kind if world > 1 else "none", "launches_per_step": lps,
                       "per_gpu_weight_bytes": by["total"],
                       "frac_of_hbm_roofline_per_gpu": (by["total"] / (peak * 1e9)) / (ms / args.steps * 1e-3)},
            "gpu_launches": lps * args.steps}), flush=True)
    if world > 1:
        dist.destroy_process_group()
