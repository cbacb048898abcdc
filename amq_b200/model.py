"""Mixed-precision decoder assembled from kernel-native quantized linears + the decode-step glue.

This is what `amq_speed_benchmark.py` builds by swapping GPTQLinear / FT_QuantLinear modules into
an HF model per the searched bit config (/root/reference/amq/amq_speed_benchmark.py:231-251) and
then times (amq/utils/speed.py:23-46, 93-125).  Here the model is random-init (no checkpoints
offline), built layer by layer directly in the kernel-native layout so fp16 weights are never
materialised (mandatory for 70B), and a decode step is a fixed chain of C-ABI launches

    [rmsnorm -> q|k|v GEMV] -> [rope + kv append + attention] -> [o_proj GEMV + residual]
    -> [rmsnorm -> gate|up GEMV] -> [silu*up -> down_proj GEMV + residual]      (x n_block)
    -> [final rmsnorm -> lm_head GEMV] -> [argmax]

captured once in a CUDA graph with programmatic dependent launch between the kernels.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops
from ._lib import PRO_MUL, PRO_NONE, PRO_RMSNORM, PRO_SILU_MUL, check, cur_stream, lib, ptr
from .arch import LINEARS, ModelShape

GROUP = 128
_SCALE_RANGE = {2: (0.024, 0.054), 3: (0.010, 0.023), 4: (0.0047, 0.011)}   # SURVEY §8d realistic ranges


def synthetic_native(bits: int, N: int, K: int, device, gen: torch.Generator) -> torch.Tensor:
    """Random codes + realistic fp16 scale / zero*scale written straight into the native layout."""
    nb = ops.native_bytes(bits, N, K)
    rec = bits * 512 + 128
    buf = torch.randint(0, 256, (nb // rec, rec), dtype=torch.uint8, device=device, generator=gen)
    lo, hi = _SCALE_RANGE[bits]
    n_meta = (nb // rec) * 32
    scale = torch.empty(n_meta, device=device, dtype=torch.float32).uniform_(lo, hi, generator=gen)
    zero = torch.empty(n_meta, device=device, dtype=torch.float32).uniform_(0.5, 2 ** bits - 1.5, generator=gen)
    # keep the dequantised weights ~N(0, 0.02): centre the zero so (q - zero) is roughly symmetric
    meta = torch.stack([scale.half(), (zero.half() * scale.half())], dim=1).contiguous()     # half2(scale, zero*scale)
    buf[:, bits * 512:] = meta.view(torch.uint8).reshape(nb // rec, 128)
    return buf.reshape(-1)


class QuantDecoder:
    """Random-init Llama / Mistral / Qwen2 decoder with per-linear bit widths (arch dict of
    App. A4: {'self_attn.q_proj': [bits per block], ...})."""

    def __init__(self, shape: ModelShape, arch: Dict[str, List[int]], batch: int = 1, max_seq: int = 256,
                 device: str = "cuda:0", seed: int = 0, n_block: Optional[int] = None, pdl: bool = True,
                 tp_rank: int = 0, tp_world: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("amq_b200.QuantDecoder needs a CUDA device (no CPU path)")
        self.shape = shape
        self.arch = arch
        self.B = batch
        self.max_seq = max_seq
        self.dev = torch.device(device)
        self.n_block = n_block if n_block is not None else shape.n_block
        self.pdl = pdl
        self.tp_rank, self.tp_world = tp_rank, tp_world
        self.H, self.I, self.D = shape.hidden, shape.inter, shape.head_dim
        tp = tp_world
        assert shape.n_heads % tp == 0 and shape.n_kv_heads % tp == 0 and shape.inter % (tp * GROUP) == 0
        self.Hq, self.Hkv = shape.n_heads // tp, shape.n_kv_heads // tp
        self.I_loc = shape.inter // tp
        self.q_dim, self.kv_dim = self.Hq * self.D, self.Hkv * self.D
        torch.cuda.set_device(self.dev)
        check(lib().amqb_preload(), "preload")          # no first-use kernel loads (context synchronisation) inside a decode step
        gen = torch.Generator(device=self.dev).manual_seed(seed)
        H, B = self.H, batch
        self.embed = (torch.randn(shape.vocab, H, device=self.dev, generator=gen) * 0.02).half()
        self.lm_head = (torch.randn(shape.vocab, H, device=self.dev, generator=gen) * 0.02).half()
        self.final_norm = torch.ones(H, device=self.dev, dtype=torch.float16)
        self.layers = []
        self.weight_bytes = 0
        for li in range(self.n_block):
            L = {}
            dims = {"self_attn.q_proj": (self.q_dim, H), "self_attn.k_proj": (self.kv_dim, H),
                    "self_attn.v_proj": (self.kv_dim, H), "self_attn.o_proj": (H, self.q_dim),
                    "mlp.gate_proj": (self.I_loc, H), "mlp.up_proj": (self.I_loc, H), "mlp.down_proj": (H, self.I_loc)}
            for name in LINEARS:
                bits = int(arch[name][li % len(arch[name])])
                N, K = dims[name]
                L[name] = (bits, synthetic_native(bits, N, K, self.dev, gen), N, K)
                self.weight_bytes += L[name][1].numel()
            L["norm1"] = torch.ones(H, device=self.dev, dtype=torch.float16)
            L["norm2"] = torch.ones(H, device=self.dev, dtype=torch.float16)
            if shape.qkv_bias:
                L["qkv_bias"] = (torch.randn(self.q_dim + 2 * self.kv_dim, device=self.dev, generator=gen) * 0.02).half()
            L["k_cache"] = torch.zeros(B, self.Hkv, max_seq, self.D, device=self.dev, dtype=torch.float16)
            L["v_cache"] = torch.zeros(B, self.Hkv, max_seq, self.D, device=self.dev, dtype=torch.float16)
            self.layers.append(L)
        # step buffers
        qkv_w = self.q_dim + 2 * self.kv_dim
        self.tokens = torch.zeros(B, dtype=torch.int64, device=self.dev)
        self.h = torch.zeros(B, H, device=self.dev, dtype=torch.float16)
        self.qkv = torch.zeros(B, qkv_w, device=self.dev, dtype=torch.float16)
        self.attn = torch.zeros(B, self.q_dim, device=self.dev, dtype=torch.float16)
        self.gu = torch.zeros(B, 2 * self.I_loc, device=self.dev, dtype=torch.float16)
        self.part = torch.zeros(B, H, device=self.dev, dtype=torch.float16)      # TP partial sums
        self.logits = torch.zeros(B, shape.vocab, device=self.dev, dtype=torch.float32)
        self.next_tokens = torch.zeros(B, dtype=torch.int64, device=self.dev)
        self.pos = torch.zeros(1, dtype=torch.int32, device=self.dev)
        # generated ids of every step since the last generate() / reset(): [steps][B], filled by the step's last launch
        self.token_log = torch.zeros(max_seq, B, dtype=torch.int64, device=self.dev)
        self.log_pos = torch.zeros(B, dtype=torch.int32, device=self.dev)
        self.rope = torch.empty(max_seq, self.D // 2, 2, dtype=torch.float32, device=self.dev)
        check(lib().amqb_rope_table(ptr(self.rope), max_seq, self.D, ctypes.c_float(shape.rope_theta), cur_stream()), "rope_table")
        self.ws = ops.workspace(self.dev, 4 * shape.inter, max(shape.inter, shape.hidden), batch)
        # long contexts: the cached positions of a head are shared by several CTAs (amqb_attn_decode_split) when one CTA
        # per (head, sequence) would leave most of the chip idle; below ATTN_SPLIT_MIN_POS it is the single-CTA kernel
        sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        self.attn_splits = max(1, min(4, sms // max(1, self.Hq * batch)))      # per rank: Hq is this rank's share of the heads
        self.attn_split_min_pos = int(os.environ.get("AMQB_ATTN_SPLIT_MIN_POS", "256"))
        self.attn_ws = torch.zeros(max(256, int(lib().amqb_attn_split_workspace_bytes(batch, self.Hq, self.D, self.attn_splits))),
                                   dtype=torch.uint8, device=self.dev)
        # Two captured steps: the short-context one launches one CTA per head (the extra CTAs of the split grid, although
        # they leave at once below the threshold, were measured to cost ~0.75 us per layer); step() picks by a host-side
        # mirror of the position.  Either graph is correct at any position — the mirror only selects the faster one.
        # SiLU in gate_proj's epilogue is a batch-1 feature of the kernels; AMQB_SILU_IN_CONSUMER=1: A/B switch (round-1 placement)
        self.silu_in_consumer = batch > 1 or os.environ.get("AMQB_SILU_IN_CONSUMER") == "1"
        self._cur_splits = 1
        self._pos_h = 0
        self.graph_long: Optional[torch.cuda.CUDAGraph] = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.allreduce = None          # set by attach_allreduce (amq_b200.tp) for tensor-parallel runs
        self.ar_ctx = None             # amqb_ar_ctx when the all-reduce is fused into the row-parallel GEMVs' epilogue
        self.ar_gen = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.launches_per_step = 0
        self._build_problems()

    # ---------------------------------------------------------------- tensor parallel
    def attach_allreduce(self, ar, fused: bool = True) -> None:
        """ar: tp.PeerAllReduce / LocalAllReduce / NcclAllReduce.  fused (peer-memory kinds only): o_proj and down_proj
        push their fp32 partial sums to the peers from their own epilogue and finish h = h + sum over ranks there
        (amqb_gemv_problem.allreduce) — no all-reduce launch; otherwise one all-reduce call follows each of them."""
        self.allreduce = ar
        self.ar_ctx = ar.make_ctx(self.pos, self.ar_gen) if (fused and hasattr(ar, "make_ctx")) else None
        self._build_problems()
        self.graph = self.graph_long = None

    def bump_generation(self) -> None:
        """Whenever positions restart (reset, the rewind after a capture warm-up) the fused all-reduce's epoch
        (generation, position, call index) must not repeat: slots still hold the earlier pass's values."""
        self.ar_gen.add_(1)

    # ---------------------------------------------------------------- tensor-parallel shard of another decoder
    def adopt_shard_of(self, full: "QuantDecoder") -> None:
        """Replace this rank's random weights by rank tp_rank's Megatron shard of `full` (a tp_world = 1 decoder of the
        same shape / arch on the same device): column-parallel linears take a run of 32-row record blocks, row-parallel
        ones a run of k groups of every row block.  Parity-test / emulation aid."""
        assert full.tp_world == 1 and full.shape == self.shape and full.n_block == self.n_block and full.B == self.B
        r, w = self.tp_rank, self.tp_world
        self.embed, self.lm_head, self.final_norm = full.embed, full.lm_head, full.final_norm
        for L, F in zip(self.layers, full.layers):
            for name in LINEARS:
                bits, buf, N, K = F[name]
                rec = bits * 512 + 128
                v = buf.view(N // 32, K // GROUP, rec)
                _, _, n_loc, k_loc = L[name]
                if name in ("self_attn.o_proj", "mlp.down_proj"):          # row-parallel: split K
                    g0 = r * (k_loc // GROUP)
                    shard = v[:, g0: g0 + k_loc // GROUP]
                else:                                                        # column-parallel: split N
                    b0 = r * (n_loc // 32)
                    shard = v[b0: b0 + n_loc // 32]
                L[name] = (bits, shard.contiguous().reshape(-1), n_loc, k_loc)
            L["norm1"], L["norm2"] = F["norm1"], F["norm2"]
            if "qkv_bias" in F:
                fq, fk = full.q_dim, full.kv_dim
                b = F["qkv_bias"]
                L["qkv_bias"] = torch.cat([b[r * self.q_dim: (r + 1) * self.q_dim],
                                           b[fq + r * self.kv_dim: fq + (r + 1) * self.kv_dim],
                                           b[fq + fk + r * self.kv_dim: fq + fk + (r + 1) * self.kv_dim]]).contiguous()
        self._build_problems()
        self.graph = self.graph_long = None

    # ---------------------------------------------------------------- static launch descriptors
    def _build_problems(self) -> None:
        eps = self.shape.rms_eps
        B = self.B
        self._plan = []
        for L in self.layers:
            qb, qw, qn, qk = L["self_attn.q_proj"]
            kb, kw, kn, kk = L["self_attn.k_proj"]
            vb, vw, vn, vk = L["self_attn.v_proj"]
            qkv_ld = self.qkv.stride(0)
            bias = L.get("qkv_bias")

            def sub(t, off):
                return ctypes.c_void_p(t.data_ptr() + 2 * off)

            def prob(bits, w, x, y_ptr, N, K, ldx, ldy, bias_ptr=None, residual=None, pro=PRO_NONE, gamma=None):
                p = ops.GemvProblem()
                p.bits, p.M, p.N, p.K = bits, B, N, K
                p.w_native = w.data_ptr()
                p.x = x.data_ptr()
                p.ldx = ldx
                p.y = y_ptr
                p.ldy = ldy
                p.bias = bias_ptr
                p.residual = residual.data_ptr() if residual is not None else None
                p.prologue = pro
                p.gamma = gamma.data_ptr() if gamma is not None else None
                p.eps = eps
                return p

            def bptr(off):
                return (bias.data_ptr() + 2 * off) if bias is not None else None

            qkv = [prob(qb, qw, self.h, self.qkv.data_ptr(), qn, qk, self.H, qkv_ld, bptr(0), None, PRO_RMSNORM, L["norm1"]),
                   prob(kb, kw, self.h, self.qkv.data_ptr() + 2 * self.q_dim, kn, kk, self.H, qkv_ld, bptr(self.q_dim), None,
                        PRO_RMSNORM, L["norm1"]),
                   prob(vb, vw, self.h, self.qkv.data_ptr() + 2 * (self.q_dim + self.kv_dim), vn, vk, self.H, qkv_ld,
                        bptr(self.q_dim + self.kv_dim), None, PRO_RMSNORM, L["norm1"])]
            ob, ow, on, ok = L["self_attn.o_proj"]
            tp = self.tp_world > 1 and self.ar_ctx is None     # partial sums to `part`, a separate all-reduce adds them to h
            li = len(self._plan)

            def fuse(p, call):
                if self.ar_ctx is not None:
                    p.allreduce = ctypes.pointer(self.ar_ctx)
                    p.ar_call = call & 0xFF
                return p

            o = [fuse(prob(ob, ow, self.attn, (self.part if tp else self.h).data_ptr(), on, ok, self.q_dim, self.H, None,
                           None if tp else self.h), 2 * li)]
            gb, gw, gn, gk = L["mlp.gate_proj"]
            ub, uw, un, uk = L["mlp.up_proj"]
            gu_ld = self.gu.stride(0)
            gu = [prob(gb, gw, self.h, self.gu.data_ptr(), gn, gk, self.H, gu_ld, None, None, PRO_RMSNORM, L["norm2"]),
                  prob(ub, uw, self.h, self.gu.data_ptr() + 2 * self.I_loc, un, uk, self.H, gu_ld, None, None, PRO_RMSNORM,
                       L["norm2"])]
            db, dw, dn, dk = L["mlp.down_proj"]
            down = [fuse(prob(db, dw, self.gu, (self.part if tp else self.h).data_ptr(), dn, dk, gu_ld, self.H, None,
                              None if tp else self.h, PRO_SILU_MUL if self.silu_in_consumer else PRO_MUL), 2 * li + 1)]
            # SiLU once per element, in gate_proj's epilogue (act), not once per element in each of down_proj's ~128 CTAs
            gu[0].act = 0 if self.silu_in_consumer else 1
            # scheduling hint (include/amqb.h, after_gemv): which launches directly follow another batch-1 GEMV launch on the
            # stream.  q|k|v follows the previous layer's down_proj, gate|up follows o_proj, down_proj follows gate|up;
            # o_proj follows the attention kernel, and a separate all-reduce kernel breaks the chain as well.
            separate_ar = self.allreduce is not None and self.ar_ctx is None
            qkv[0].after_gemv = int(li > 0 and not separate_ar)
            gu[0].after_gemv = int(not separate_ar)
            down[0].after_gemv = 1
            self._plan.append({"qkv": (ops.GemvProblem * 3)(*qkv), "o": (ops.GemvProblem * 1)(*o),
                               "gu": (ops.GemvProblem * 2)(*gu), "down": (ops.GemvProblem * 1)(*down), "L": L})

    # ---------------------------------------------------------------- one decode step (all launches)
    def _gemv(self, arr, n) -> None:
        check(lib().amqb_gemv_grouped(arr, n, ptr(self.ws), ctypes.c_size_t(self.ws.numel()), int(self.pdl), cur_stream()),
              "gemv_grouped")
        self.launches_per_step += len({arr[i].prologue for i in range(n)})

    def _step_launches(self) -> None:
        Lb = lib()
        st = cur_stream()
        S = self.shape
        self.launches_per_step = 0
        check(Lb.amqb_embed(ptr(self.tokens), ptr(self.embed), ptr(self.h), self.B, self.H, st), "embed")
        self.launches_per_step += 1
        for P in self._plan:
            L = P["L"]
            self._gemv(P["qkv"], 3)
            check(Lb.amqb_attn_decode_split(ptr(self.qkv), ptr(L["k_cache"]), ptr(L["v_cache"]), ptr(self.attn), ptr(self.pos),
                                            self.B, self.Hq, self.Hkv, self.D, self.max_seq, ctypes.c_float(S.rope_theta),
                                            ptr(self.rope), self._cur_splits, self.attn_split_min_pos, ptr(self.attn_ws),
                                            ctypes.c_size_t(self.attn_ws.numel()), st), "attn_decode_split")
            self.launches_per_step += 1
            self._gemv(P["o"], 1)
            if self.allreduce is not None and self.ar_ctx is None:
                self.allreduce(self.part, self.h)          # h += sum over ranks of part
                self.launches_per_step += 1
            self._gemv(P["gu"], 2)
            self._gemv(P["down"], 1)
            if self.allreduce is not None and self.ar_ctx is None:
                self.allreduce(self.part, self.h)
                self.launches_per_step += 1
        check(Lb.amqb_lm_head(ptr(self.lm_head), ptr(self.h), ptr(self.final_norm), ctypes.c_float(S.rms_eps),
                              ptr(self.logits), self.B, S.vocab, self.H, st), "lm_head")
        # greedy token, fed back as the next input, position advanced: all in the step's last launch
        check(Lb.amqb_argmax_advance_log(ptr(self.logits), ptr(self.next_tokens), ptr(self.tokens), ptr(self.pos),
                                         ptr(self.token_log), self.token_log.shape[0], ptr(self.log_pos), self.B, S.vocab, st),
              "argmax_advance")
        self.launches_per_step += 2

    def _long_context(self) -> bool:
        return self.attn_splits > 1 and self._pos_h >= self.attn_split_min_pos

    def _check_room(self, steps: int = 1) -> None:
        """The attention kernels index the K/V cache and the rope table straight from the device-side position: refuse a
        step that would run past max_seq instead of writing past the cache."""
        if self._pos_h + steps > self.max_seq:
            raise RuntimeError(f"QuantDecoder: position {self._pos_h} + {steps} step(s) exceeds max_seq {self.max_seq}")

    def _capture(self, host_in: Optional[torch.Tensor] = None, host_out: Optional[torch.Tensor] = None, splits: int = 1,
                 stream: Optional[torch.cuda.Stream] = None, warm: bool = True):
        """warm=False: the caller has already run the step eagerly (tp.LocalTPGroup warms all emulated ranks together:
        a rank's step cannot finish before its peers' steps have been launched)."""
        self._check_room(2)                         # the warm-up below runs two real steps from the current position
        lib().amqb_set_pdl(int(self.pdl))
        self._cur_splits = splits
        s = stream if stream is not None else torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            if warm:
                saved_pos, saved_tok, saved_log = self.pos.clone(), self.tokens.clone(), self.log_pos.clone()
                for _ in range(2):                  # warm-up outside capture (lazy module loads, attributes)
                    self._step_launches()
                self.pos.copy_(saved_pos)           # the step advances position and input ids itself: undo the warm-up's
                self.tokens.copy_(saved_tok)
                self.log_pos.copy_(saved_log)
                self.bump_generation()
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                if host_in is not None:
                    self.tokens.copy_(host_in, non_blocking=True)      # memcpy node: pinned host -> device
                self._step_launches()
                if host_out is not None:
                    host_out.copy_(self.tokens, non_blocking=True)     # memcpy node: device -> pinned host
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self._cur_splits = 1
        return g

    def capture(self) -> None:
        """Capture one decode step (+ feeding the argmax back as the next input, + position advance)
        into a CUDA graph; when the cache can grow past the split threshold, a second one with the split attention grid."""
        self.graph = self._capture()
        if self.attn_splits > 1 and self.max_seq > self.attn_split_min_pos:
            self.graph_long = self._capture(splits=self.attn_splits)

    def step(self) -> None:
        """One token for every sequence of the batch: replay the captured graph."""
        if self.graph is None:
            self.capture()
        self._check_room()
        (self.graph_long if (self.graph_long is not None and self._long_context()) else self.graph).replay()
        self._pos_h += 1

    def step_host(self, host_in: torch.Tensor, host_out: torch.Tensor) -> None:
        """Host-facing step: this step's input ids are read from the pinned host tensor `host_in` [B] and the
        generated ids land in the pinned host tensor `host_out` [B] before the call returns.  Both copies are
        memcpy nodes of the captured graph (one launch + one synchronise per token on the host side)."""
        if not (host_in.is_pinned() and host_out.is_pinned()):
            raise ValueError("step_host: host_in / host_out must be pinned host tensors")
        if host_in.shape != self.tokens.shape or host_out.shape != self.tokens.shape or \
                host_in.dtype != self.tokens.dtype or host_out.dtype != self.tokens.dtype:
            raise ValueError("step_host: host tensors must match the token buffer [B] int64")
        key = (host_in.data_ptr(), host_out.data_ptr())
        if getattr(self, "_io_key", None) != key:
            self._io_graphs, self._io_key = {False: self._capture(host_in, host_out)}, key
            if self.attn_splits > 1 and self.max_seq > self.attn_split_min_pos:
                self._io_graphs[True] = self._capture(host_in, host_out, splits=self.attn_splits)
            self._io_bufs = (host_in, host_out)                         # keep the captured addresses alive
        self._check_room()
        self._io_graphs[self._long_context() and True in self._io_graphs].replay()
        self._pos_h += 1
        torch.cuda.current_stream(self.dev).synchronize()

    def step_eager(self) -> None:
        self._check_room()
        lib().amqb_set_pdl(0)
        saved = self.pdl
        self.pdl = False
        self._cur_splits = self.attn_splits if self._long_context() else 1
        self._step_launches()
        self._cur_splits = 1
        self._pos_h += 1
        self.pdl = saved

    def reset(self) -> None:
        self.pos.zero_()
        self.log_pos.zero_()
        self._pos_h = 0
        self.bump_generation()

    # ---------------------------------------------------------------- prompt prefill (all prompt rows at once)
    @torch.inference_mode()
    def prefill(self, ids: torch.Tensor, use_graph: bool = True) -> None:
        """Consume ids [B, T] at cache positions pos .. pos+T-1 in ONE pass over the weights: every linear runs once
        on the [B*T, K] activation matrix (tcgen05 GEMM for more than 16 rows, the skinny decode kernel below,
        ops.linear_forward — the reference modules' own row-count dispatch, autogptq.py:163, ft.py:129-142), with the
        row kernels of csrc/prefill_glue.cu in between.  Leaves the K/V cache filled and the position advanced; the
        hidden state of the prompt rows is not kept (the last layer stops after the cache append), so the caller feeds
        the LAST prompt token through step() to obtain the first logits.  What HF generate() does with the prompt in
        the reference's benchmark_tps (speed.py:23-46).  The ~14 launches per layer are captured in a CUDA graph per
        (T, start position) and replayed (a 64-token prompt is otherwise bound by the host's launch rate).
        Tensor-parallel decoders run the same pass over their shard: the partial sums of o_proj / down_proj go through
        the attached all-reduce's rows() (tp.PeerAllReduce(..., rows_elems=...): amqb_allreduce_rows_f16); every rank
        must call prefill() with the same ids, and the start position is taken from the host mirror of the position."""
        if self.tp_world > 1 and not getattr(self.allreduce, "has_rows", False):
            raise RuntimeError("prefill: a tensor-parallel decoder needs an all-reduce with a rows() method attached "
                               "(tp.PeerAllReduce(..., rows_elems=B * max_seq * hidden), tp.NcclAllReduce)")
        ids = ids.to(self.dev).to(torch.int64).contiguous()
        B, T = ids.shape
        assert B == self.B and T >= 1
        # tensor parallel: the host mirror (reading the device word would wait for this rank's stream, which in the
        # one-process emulation of tp.LocalTPGroup may hold launches that wait for a peer not yet launched)
        pos0 = self._pos_h if self.tp_world > 1 else int(self.pos.item())
        assert pos0 + T <= self.max_seq
        self._pos_h = pos0 + T
        if not use_graph:
            self._prefill_launches(ids, pos0)
            self.pos.add_(T)
            return
        if not hasattr(self, "_pf_graphs"):
            self._pf_graphs = {}
        key = (T, pos0)
        if key not in self._pf_graphs:
            if len(self._pf_graphs) >= 4:                     # each graph pins its activation pool: keep a few shapes
                self._pf_graphs.pop(next(iter(self._pf_graphs)))
            static_ids = ids.clone()
            cur = torch.cuda.current_stream(self.dev)
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                self._prefill_launches(static_ids, pos0)       # warm-up outside capture (workspaces, attributes)
                s.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    self._prefill_launches(static_ids, pos0)
            cur.wait_stream(s)
            self._pf_graphs[key] = (g, static_ids)
        g, static_ids = self._pf_graphs[key]
        static_ids.copy_(ids)
        g.replay()
        self.pos.add_(T)

    def _prefill_launches(self, ids: torch.Tensor, pos0: int) -> None:
        Lb, st = lib(), cur_stream()
        S = self.shape
        B, T = ids.shape
        M, H = B * T, self.H
        f16 = dict(dtype=torch.float16, device=self.dev)
        h = torch.empty(M, H, **f16)
        x = torch.empty(M, H, **f16)
        attn = torch.empty(M, self.q_dim, **f16)
        act = torch.empty(M, self.I_loc, **f16)
        check(Lb.amqb_embed(ptr(ids.reshape(-1)), ptr(self.embed), ptr(h), M, H, st), "embed")

        def linear(L, name, inp, bias=None):
            bits, w, N, K = L[name]
            return ops.linear_forward(bits, w, inp, N, K, bias)

        def add_norm(y, gamma):
            """h += y, x = RMSNorm(h) * gamma (what follows every o_proj / down_proj).  One GPU: one fused row kernel.
            Tensor parallel: y holds this rank's partial sums of a row-parallel linear (its K shard), so the add is the
            all-reduce h += sum over ranks of y, in rank order (bit-identical on every rank), SURVEY §8e."""
            if self.tp_world > 1:
                self.allreduce.rows(y, h)
                check(Lb.amqb_rmsnorm_rows(ptr(h), ptr(gamma), ctypes.c_float(S.rms_eps), ptr(x), M, H, st), "rmsnorm_rows")
            else:
                check(Lb.amqb_add_rmsnorm_rows(ptr(h), ptr(y), ptr(gamma), ctypes.c_float(S.rms_eps), ptr(x), M, H, st),
                      "add_rmsnorm_rows")

        check(Lb.amqb_rmsnorm_rows(ptr(h), ptr(self.layers[0]["norm1"]), ctypes.c_float(S.rms_eps), ptr(x), M, H, st), "rmsnorm_rows")
        for li, L in enumerate(self.layers):
            bias = L.get("qkv_bias")
            bq = bk = bv = None
            if bias is not None:
                bq, bk, bv = bias[: self.q_dim], bias[self.q_dim: self.q_dim + self.kv_dim], bias[self.q_dim + self.kv_dim:]
            # q|k|v (and gate|up below) share their activations: one grouped launch each
            q, k, v = ops.linear_forward_grouped(
                [(L[n][0], L[n][1], L[n][2], b_) for n, b_ in (("self_attn.q_proj", bq), ("self_attn.k_proj", bk), ("self_attn.v_proj", bv))],
                x, H)
            check(Lb.amqb_attn_prefill(ptr(q), ptr(k), ptr(v), ptr(L["k_cache"]), ptr(L["v_cache"]), ptr(attn), pos0, T, B,
                                       self.Hq, self.Hkv, self.D, self.max_seq, ptr(self.rope), st), "attn_prefill")
            if li == len(self.layers) - 1:
                break                                   # only the cache rows of the last layer are needed
            o = linear(L, "self_attn.o_proj", attn)
            add_norm(o, L["norm2"])
            g, u = ops.linear_forward_grouped([(L[n][0], L[n][1], L[n][2], None) for n in ("mlp.gate_proj", "mlp.up_proj")], x, H)
            check(Lb.amqb_silu_mul_rows(ptr(g), ptr(u), ptr(act), M, self.I_loc, st), "silu_mul_rows")
            d = linear(L, "mlp.down_proj", act)
            add_norm(d, self.layers[li + 1]["norm1"])   # the next layer's input norm

    @torch.inference_mode()
    def generate(self, input_ids: torch.Tensor, max_new_tokens: int, use_graph: bool = True,
                 prefill: Optional[bool] = None) -> torch.Tensor:
        """Greedy generation like model.generate(do_sample=False, min=max_new_tokens) in the reference's
        benchmark_tps (speed.py:23-46).  input_ids: [B, prompt] (host or device).  prefill=True (default on one
        GPU): prompt tokens 0..P-2 go through prefill() in one pass, the last prompt token through the decode
        step; prefill=False: the whole prompt is consumed token by token through the decode path."""
        assert input_ids.shape[0] == self.B
        prompt = input_ids.shape[1]
        assert prompt + max_new_tokens <= self.max_seq
        ids = input_ids.to(self.dev, non_blocking=True)
        self.reset()
        fn = self.step if use_graph else self.step_eager
        if use_graph and self.graph is None:
            self.capture()
        if prefill is None:
            prefill = self.tp_world == 1 or getattr(self.allreduce, "has_rows", False)
        first = 0
        if prefill and prompt > 1:
            self.prefill(ids[:, : prompt - 1])
            first = prompt - 1
        for t in range(first, prompt):
            self.tokens.copy_(ids[:, t])
            if t == prompt - 1:
                self.log_pos.zero_()                 # the step fed with the last prompt token logs generated id 0
            fn()
        # the generated ids are logged on the device by each step's last launch (amqb_argmax_advance_log): the loop is
        # graph replays back to back, with no per-token copy kernel between them
        for t in range(1, max_new_tokens):
            fn()
        return self.token_log[:max_new_tokens].t().contiguous()

    # ---------------------------------------------------------------- accounting
    def algorithmic_bytes_per_token(self) -> Dict[str, int]:
        lin = 0
        for L in self.layers:
            for name in LINEARS:
                bits, _, N, K = L[name]
                lin += N * K * bits // 8 + (K // GROUP) * N * 4
        head = 2 * self.shape.vocab * self.H
        return {"linears": lin, "lm_head": head, "total": lin + head}
