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

    def __init__(self, rank: int, world: int, max_elems: int, pdl: bool = True, rows_elems: int = 0):
        """COLLECTIVE (handles travel over torch.distributed).  rows_elems > 0 also creates the exchange buffers of the
        prompt pass's [B*T, hidden] all-reduce (amqb_allreduce_rows_f16): rows_elems >= B * T_max * hidden."""
        from ._lib import lib
        self.rank, self.world, self.max_elems, self.pdl = rank, world, max_elems, pdl
        self.rows_elems = rows_elems
        self.has_rows = rows_elems > 0
        self._opened = []
        self._mine, self._peers = self._exchange(int(lib().amqb_ar_buffer_bytes(max_elems, world)))
        self._rows_mine = self._rows_peers = None
        if rows_elems > 0:
            self._rows_mine, self._rows_peers = self._exchange(int(lib().amqb_ar_rows_buffer_bytes(rows_elems, world)))
        import torch.distributed as dist
        dist.barrier()

    def _exchange(self, nbytes: int):
        from ._lib import check, lib
        L = lib()
        mine = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        check(L.amqb_ar_alloc(ctypes.c_size_t(nbytes), ctypes.byref(mine), handle), "ar_alloc")
        handles = exchange_handles(bytes(handle))
        peers = (ctypes.c_void_p * self.world)()
        for r, h in enumerate(handles):
            if r == self.rank:
                peers[r] = mine
            else:
                p = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                check(L.amqb_ar_open(buf, ctypes.byref(p)), "ar_open")
                peers[r] = p
                self._opened.append(p)
        return mine, peers

    def rows(self, partial: torch.Tensor, h: torch.Tensor) -> None:
        """h <- h + sum over ranks of partial for an [M, hidden] matrix (the prompt pass)."""
        if self._rows_peers is None:
            raise RuntimeError("PeerAllReduce: created without rows_elems (no exchange buffers for the prompt pass)")
        _rows_call(self._rows_peers, self.rank, self.world, partial, h, self.rows_elems, self.pdl)

    def make_ctx(self, pos_dev: torch.Tensor, gen_dev: torch.Tensor):
        """amqb_ar_ctx for the all-reduce fused into the row-parallel GEMV's epilogue (amqb_gemv_problem.allreduce)."""
        return _make_ctx(self._peers, self.rank, self.world, self.max_elems, pos_dev, gen_dev)

    def close(self) -> None:
        """COLLECTIVE: every rank unmaps its peers' buffers, then (after a barrier: nobody may free a buffer a peer
        still has mapped) frees its own.  Never called implicitly — a destructor would run at a different time on
        every rank."""
        import torch.distributed as dist
        from ._lib import lib
        if getattr(self, "_mine", None) is None:
            return
        torch.cuda.synchronize()
        L = lib()
        for p in self._opened:
            L.amqb_ar_close(p)
        self._opened = []
        if dist.is_initialized():
            dist.barrier()
        L.amqb_ar_free(self._mine)
        if self._rows_mine is not None:
            L.amqb_ar_free(self._rows_mine)
        self._mine = self._rows_mine = self._rows_peers = None

    def __call__(self, partial: torch.Tensor, h: torch.Tensor) -> None:
        from ._lib import check, cur_stream, lib, ptr
        check(lib().amqb_allreduce_f16(self._peers, self.rank, self.world, ptr(partial), ptr(h), ptr(h), partial.numel(),
                                       self.max_elems, int(self.pdl), cur_stream()), "allreduce")


def _rows_call(peers, rank: int, world: int, partial: torch.Tensor, h: torch.Tensor, max_elems: int, pdl: bool) -> None:
    from ._lib import check, cur_stream, lib, ptr
    if partial.dtype != torch.float16 or h.dtype != torch.float16 or partial.numel() != h.numel():
        raise ValueError("all-reduce of rows: fp16 partial and residual of the same size")
    check(lib().amqb_allreduce_rows_f16(peers, rank, world, ptr(partial), ptr(h), ptr(h), ctypes.c_longlong(partial.numel()),
                                        ctypes.c_longlong(max_elems), int(pdl), cur_stream()), "allreduce_rows")


def _make_ctx(peers, rank: int, world: int, max_elems: int, pos_dev: torch.Tensor, gen_dev: torch.Tensor):
    from ._lib import ArCtx
    c = ArCtx()
    for r in range(world):
        c.peer_bufs[r] = peers[r]
    c.rank, c.world, c.max_elems = rank, world, max_elems
    c.pos_dev, c.gen_dev = pos_dev.data_ptr(), gen_dev.data_ptr()
    return c


class LocalAllReduce:
    """PeerAllReduce for emulated ranks that live in ONE process on ONE device (tp.LocalTPGroup): the exchange buffers
    are plain device tensors of this process, so the "peer" pointers need no IPC mapping.  Same kernel."""

    def __init__(self, rank: int, world: int, max_elems: int, bufs: List[torch.Tensor], pdl: bool = True,
                 rows_elems: int = 0, rows_bufs: Optional[List[torch.Tensor]] = None):
        self.rank, self.world, self.max_elems, self.pdl = rank, world, max_elems, pdl
        self._bufs = bufs
        self._peers = (ctypes.c_void_p * world)(*[ctypes.c_void_p(b.data_ptr()) for b in bufs])
        self.rows_elems, self._rows_bufs = rows_elems, rows_bufs
        self.has_rows = bool(rows_bufs)
        self._rows_peers = (ctypes.c_void_p * world)(*[ctypes.c_void_p(b.data_ptr()) for b in rows_bufs]) if rows_bufs else None

    def rows(self, partial: torch.Tensor, h: torch.Tensor) -> None:
        if self._rows_peers is None:
            raise RuntimeError("LocalAllReduce: created without rows_bufs (no exchange buffers for the prompt pass)")
        _rows_call(self._rows_peers, self.rank, self.world, partial, h, self.rows_elems, self.pdl)

    def make_ctx(self, pos_dev: torch.Tensor, gen_dev: torch.Tensor):
        return _make_ctx(self._peers, self.rank, self.world, self.max_elems, pos_dev, gen_dev)

    def __call__(self, partial: torch.Tensor, h: torch.Tensor) -> None:
        from ._lib import check, cur_stream, lib, ptr
        check(lib().amqb_allreduce_f16(self._peers, self.rank, self.world, ptr(partial), ptr(h), ptr(h), partial.numel(),
                                       self.max_elems, int(self.pdl), cur_stream()), "allreduce")


class LocalTPGroup:
    """`world` tensor-parallel ranks emulated in one process on one device, one stream per rank: every rank runs its
    own captured decode step (its shard of the weights, its heads' K/V cache) and the all-reduces meet through device
    memory exactly as they do through NVLink-mapped peer memory.  What it is for: model-level parity of the sharded
    decoder on a single-GPU box (tests/test_gpu_tp.py) — the driver's test box has one GPU."""

    def __init__(self, full, world: int, max_seq: Optional[int] = None, fused: bool = True):
        from ._lib import lib
        from .model import QuantDecoder
        # No programmatic dependent launch between the emulated ranks' kernels: a dependent GEMV launched early sits on
        # its SM (222 KB of shared memory) waiting for an all-reduce that waits for a PEER rank's kernels, which then
        # find no free SM on the one shared device.  Real ranks own a GPU each and keep PDL on.
        pdl = False
        self.full, self.world = full, world
        dev = full.dev
        shard_plan(full.shape, world)
        nbytes = int(lib().amqb_ar_buffer_bytes(full.shape.hidden * full.B, world))
        self._bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
        # exchange buffers of the prompt pass's [B*T, hidden] all-reduce (amqb_allreduce_rows_f16)
        rows_elems = full.shape.hidden * full.B * (max_seq or full.max_seq)
        rbytes = int(lib().amqb_ar_rows_buffer_bytes(rows_elems, world))
        self._rows_bufs = [torch.zeros(rbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
        self.ranks = []
        for r in range(world):
            m = QuantDecoder(full.shape, full.arch, batch=full.B, max_seq=max_seq or full.max_seq, device=str(dev), seed=0,
                             n_block=full.n_block, pdl=pdl, tp_rank=r, tp_world=world)
            m.adopt_shard_of(full)
            # own workspace (the M > 1 pre-pass writes the integer activations there; ops.workspace is per stream)
            m.ws = torch.zeros_like(m.ws)
            m.attach_allreduce(LocalAllReduce(r, world, full.shape.hidden * full.B, self._bufs, pdl=pdl,
                                              rows_elems=rows_elems, rows_bufs=self._rows_bufs), fused=fused)
            self.ranks.append(m)
        self._captured = False
        # a GEMV whose epilogue waits for the peers' partial sums (fused all-reduce) keeps its SMs: every emulated rank
        # sizes its grids for 1 / world of the device so that all ranks' kernels can be resident at once
        self._sm_limit = max(1, torch.cuda.get_device_properties(dev).multi_processor_count // world) if fused else 0

    def _limit(self, on: bool) -> None:
        from ._lib import lib
        lib().amqb_debug_set_sm_limit(self._sm_limit if on else 0)

    def timeouts(self) -> int:
        """Flag waits that gave up (0 in a correct run)."""
        from ._lib import check, lib
        n = 0
        for b in self._bufs:
            c = ctypes.c_int(0)
            check(lib().amqb_ar_timeouts(ctypes.c_void_p(b.data_ptr()), ctypes.byref(c)), "ar_timeouts")
            n += c.value
        for b in self._rows_bufs:
            c = ctypes.c_int(0)
            check(lib().amqb_ar_rows_timeouts(ctypes.c_void_p(b.data_ptr()), ctypes.byref(c)), "ar_rows_timeouts")
            n += c.value
        return n

    def set_tokens(self, tok: torch.Tensor) -> None:
        for m in self.ranks:
            m.reset()
            m.tokens.copy_(tok)
        torch.cuda.synchronize()

    def _each(self, fn) -> None:
        cur = torch.cuda.current_stream()
        for s, m in zip(self.streams, self.ranks):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                fn(m)
        for s in self.streams:
            cur.wait_stream(s)

    def step_eager(self) -> None:
        """All ranks' launches are issued (asynchronously) before anything is waited for: a rank's all-reduce spins
        until its peers' partial sums arrive."""
        def go(m):
            m.step_eager()
        self._limit(True)
        self._each(go)
        self._limit(False)

    def capture(self) -> None:
        saved = [(m.pos.clone(), m.tokens.clone(), m._pos_h) for m in self.ranks]
        for _ in range(2):
            self.step_eager()
        torch.cuda.synchronize()
        for m, (p, t, ph) in zip(self.ranks, saved):
            m.pos.copy_(p)
            m.tokens.copy_(t)
            m._pos_h = ph
            m.bump_generation()                 # positions were rewound: stale all-reduce slots must not match
        torch.cuda.synchronize()
        self._limit(True)
        for s, m in zip(self.streams, self.ranks):
            m.graph = m._capture(stream=s, warm=False)
            if m.attn_splits > 1 and m.max_seq > m.attn_split_min_pos:      # as QuantDecoder.capture
                m.graph_long = m._capture(stream=s, warm=False, splits=m.attn_splits)
        self._limit(False)
        self._captured = True

    def step(self) -> None:
        if not self._captured:
            self.capture()

        def go(m):
            m._check_room()
            (m.graph_long if (m.graph_long is not None and m._long_context()) else m.graph).replay()
            m._pos_h += 1
        self._each(go)

    def prefill(self, ids: torch.Tensor) -> None:
        """Every rank's prompt pass (QuantDecoder.prefill, plain launches: a per-rank graph capture would have to warm up
        one rank alone, whose all-reduces wait for peers that have not been launched)."""
        def go(m):
            m.prefill(ids, use_graph=False)
        self._each(go)


class NcclAllReduce:
    """Baseline: ncclAllReduce on the compute stream + residual add."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist

    def __call__(self, partial: torch.Tensor, h: torch.Tensor) -> None:
        self.dist.all_reduce(partial)
        h.add_(partial)

    rows = __call__
    has_rows = True


PROMPT_ROWS = 63          # the reference's TTFT / TPS protocol feeds a 64-token prompt: 63 rows in one pass + one decode step


def _peak_gbs() -> float:
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"])
    return 6650.0


def measure_tp70b(steps: int, warmup: int, world: int, rank: int, local: int, tp: Optional[int] = None) -> Optional[dict]:
    """Config 5: Llama-2-70B random-init, AMQ avg 3.0 bits, batch-1 decode, tensor-parallel over `tp` ranks of the
    already initialised process group (tp = 1: rank 0 alone, no collective).  CUDA-graph replay, CUDA events, max over
    the participating ranks.  Returns the record on rank 0, None elsewhere."""
    import torch.distributed as dist
    from .model import QuantDecoder
    tp = world if tp is None else tp
    shape = MODELS["Llama-2-70b-hf"]
    arch = sample_arch(shape, 3.0, seed=0)
    shard_plan(shape, tp)
    n_block = int(os.environ.get("AMQB_BLOCKS", shape.n_block))
    # fused: the all-reduce is part of the row-parallel GEMV launch (default); amqb: separate one-shot push kernel;
    # nccl: ncclAllReduce + residual add (baseline)
    ar_kind = os.environ.get("AMQB_AR", "fused") if tp > 1 else "none"
    ms, lps, by = 0.0, 0, None
    if rank < tp:
        model = QuantDecoder(shape, arch, batch=1, max_seq=max(256, warmup + steps + 8), device=f"cuda:{local}",
                             seed=0, n_block=n_block, tp_rank=rank, tp_world=tp)
        if tp > 1:
            if ar_kind == "nccl":
                model.attach_allreduce(NcclAllReduce(), fused=False)
            else:
                model.attach_allreduce(PeerAllReduce(rank, tp, shape.hidden, rows_elems=shape.hidden * PROMPT_ROWS),
                                       fused=(ar_kind == "fused"))
        model.capture()
        lps = model.launches_per_step
        by = model.algorithmic_bytes_per_token()
        model.reset()
        model.tokens.fill_(1)
        for _ in range(warmup):
            model.step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank < tp:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            model.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    # the prompt pass of the reference's protocol (speed.py:23-46: 64 prompt tokens, the first 63 in ONE pass over this
    # rank's weight shard, [63, hidden] all-reduce after o_proj / down_proj) and the decode step that yields the first token
    pf_ms = ttft_ms = 0.0
    pf_err = None
    if rank < tp:
        try:
            ids = torch.randint(0, shape.vocab, (1, PROMPT_ROWS), device=f"cuda:{local}",
                                generator=torch.Generator(device=f"cuda:{local}").manual_seed(1))
            reps = 5
            for it in range(2 + reps):             # the first call captures the pass, the second is a warm replay
                model.reset()
                model.tokens.fill_(1)
                torch.cuda.synchronize()
                a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                a.record()
                model.prefill(ids)
                b.record()
                model.step()
                c.record()
                torch.cuda.synchronize()
                if it >= 2:
                    pf_ms += a.elapsed_time(b) / reps
                    ttft_ms += a.elapsed_time(c) / reps
        except Exception as e:                     # the decode record above must survive a failure of this extra
            pf_err, pf_ms, ttft_ms = repr(e)[:200], 0.0, 0.0
    if world > 1:
        dist.barrier()
        t = torch.tensor([ms, pf_ms, ttft_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, pf_ms, ttft_ms = float(t[0]), float(t[1]), float(t[2])
    if rank < tp:
        if tp > 1 and ar_kind != "nccl":
            model.allreduce.close()             # collective over the tp ranks (tp == world here)
        del model
        torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak = _peak_gbs()
    return {"tok_s": steps / (ms * 1e-3), "ms_per_step": ms / steps, "tp": tp, "steps": steps, "warmup": warmup,
            "model": f"{shape.name} random-init ({n_block} blocks), AMQ avg 3.0 bits, batch-1 decode",
            "allreduce": ar_kind, "allreduces_per_step": 2 * n_block if tp > 1 else 0,
            "allreduce_algorithmic_bytes": 2 * shape.hidden, "launches_per_step": lps,
            "per_gpu_weight_bytes": by["total"],
            "prompt_pass_ms": pf_ms, "ttft_ms": ttft_ms, "prompt_rows": PROMPT_ROWS, "prompt_pass_error": pf_err,
            "prompt_allreduce_algorithmic_bytes": 2 * shape.hidden * PROMPT_ROWS if tp > 1 else 0,
            "frac_of_hbm_roofline_per_gpu": (by["total"] / (peak * 1e9)) / (ms / steps * 1e-3)}


def nccl_log_summary(path_glob: str) -> dict:
    """What NCCL itself logged about the communicator (NCCL_DEBUG=INFO redirected to NCCL_DEBUG_FILE by bench.py)."""
    import glob
    import re
    out = {"comm_nranks": None, "nvls": False, "p2p": False}
    for fn in glob.glob(path_glob):
        try:
            txt = open(fn, errors="replace").read()
        except OSError:
            continue
        m = re.search(r"nranks (\d+)", txt)
        if m:
            out["comm_nranks"] = max(int(m.group(1)), out["comm_nranks"] or 0)
        out["nvls"] = out["nvls"] or ("NVLS" in txt)
        out["p2p"] = out["p2p"] or ("P2P" in txt or "via P2P" in txt)
    return out


def bench_main(args) -> None:
    """bench.py --workload llama70b-tp: config 5 alone, tensor-parallel over --gpus ranks (strong scaling: same model,
    1/tp of the weights per GPU)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rec = measure_tp70b(args.steps, args.warmup, world, rank, local)
    if rank == 0:
        print(json.dumps({
            "metric": "batch-1 decode tok/s, Llama-2 70B AMQ 3-bit avg, tensor parallel", "value": rec["tok_s"], "unit": "tok/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16 (fp32 accumulate)",
            "data": "synthetic", "config": {"workload": rec["model"] + f", tp={world}", **{k: rec[k] for k in (
                "allreduce", "launches_per_step", "per_gpu_weight_bytes", "frac_of_hbm_roofline_per_gpu")}},
            "gpu_launches": rec["launches_per_step"] * args.steps}), flush=True)
    if world > 1:
        dist.destroy_process_group()
