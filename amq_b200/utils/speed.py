"""The reference's speed protocol (/root/reference/amq/utils/speed.py:15-255) on a QuantDecoder.

Same entry point, modes, argument meaning and result dictionary as the reference's `benchmark_speed`:

    benchmark_speed(model, tokenizer=None, use_ft=True, iteration=1, sizes=(1, 128, 128), mode='TPS',
                    get_peak_memory=True) -> {mode.lower(): {'B.S.G': value}, 'peak_memory': {'B.S.G': GiB}}

* TPS   gen / median wall clock of greedy generation INCLUDING the prompt pass      (speed.py:23-46)
* GeMM  1 / median wall clock of one forward over the whole prompt                 (speed.py:49-127, mode 'gemm')
* GeMV  1 / median wall clock of one synchronised single-token forward after it    (speed.py:49-127, mode 'gemv')
* TTFT  median milliseconds from prompt ids to the first generated id               (speed.py:189-237)

`model` is an amq_b200.model.QuantDecoder (the mixed-precision decoder amq_speed_benchmark.py assembles); the
wall clock brackets the same work with torch.cuda.synchronize() on both sides, as the reference does.  `use_ft`
is accepted for signature compatibility (there is one attention implementation here).
"""
from __future__ import annotations

import gc
import time
from typing import Dict, Sequence

import numpy as np
import torch

MODES = ("tps", "gemv", "gemm", "ttft")


def cleanup() -> None:
    torch.cuda.empty_cache()
    gc.collect()


@torch.inference_mode()
def forward_prompt(model, input_ids: torch.Tensor) -> torch.Tensor:
    """One forward over the whole prompt from an empty cache: rows 0..S-2 in one prefill pass, the last row through
    the decode step (whose launch chain ends in lm_head + argmax).  Returns the greedy next ids [B] (device)."""
    model.reset()
    S = input_ids.shape[1]
    if S > 1:
        if model.tp_world == 1:
            model.prefill(input_ids[:, : S - 1])
        else:
            for t in range(S - 1):
                model.tokens.copy_(input_ids[:, t])
                model.step()
    model.tokens.copy_(input_ids[:, S - 1])
    model.step()
    return model.tokens


@torch.inference_mode()
def benchmark_tps(model, input_ids: torch.Tensor, gen_seq_len: int, iteration: int) -> float:
    times = []
    for _ in range(iteration):
        cleanup()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.generate(input_ids, gen_seq_len)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    return gen_seq_len / float(np.median(times))


@torch.inference_mode()
def benchmark_gemv_gemm(model, input_ids: torch.Tensor, gen_seq_len: int, iteration: int, mode: str = "gemv") -> float:
    times = []
    for _ in range(iteration):
        cleanup()
        if mode == "gemm":
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        forward_prompt(model, input_ids)
        if mode == "gemm":
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        else:
            for _ in range(gen_seq_len):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                model.step()                       # the generated id is fed back on the device
                torch.cuda.synchronize()
                times.append(time.perf_counter() - t0)
    return 1.0 / float(np.median(times))


@torch.inference_mode()
def benchmark_speed(model, tokenizer=None, use_ft: bool = True, iteration: int = 1, sizes: Sequence[int] = (1, 128, 128),
                    mode: str = "TPS", get_peak_memory: bool = True) -> Dict[str, Dict[str, float]]:
    assert mode.lower() in MODES, "speed benchmark mode should be one of ['TPS', 'GeMV', 'GeMM', 'TTFT']"
    m = mode.lower()
    batch_size, input_seq_len, gen_seq_len = (int(s) for s in sizes)
    if batch_size != model.B:
        raise ValueError(f"benchmark_speed: sizes[0] = {batch_size} but the decoder was built for batch {model.B}")
    if input_seq_len + gen_seq_len + 1 > model.max_seq:
        raise ValueError(f"benchmark_speed: prompt {input_seq_len} + gen {gen_seq_len} exceeds the decoder's static KV cache "
                         f"({model.max_seq} positions)")
    device = model.dev
    vocab_size = model.shape.vocab
    data: Dict[str, Dict[str, float]] = {m: {}}
    if get_peak_memory:
        cleanup()
        torch.cuda.reset_peak_memory_stats(device=device)
        data["peak_memory"] = {}
    input_ids = torch.randint(0, vocab_size - 1, (batch_size, input_seq_len), dtype=torch.long).to(device)
    if model.graph is None:
        model.capture()
    model.generate(input_ids, 2)                   # device warm-up: module loads, allocator pools, graph instantiation
    cleanup()
    if get_peak_memory:
        torch.cuda.reset_peak_memory_stats(device=device)
    if m == "tps":
        speed = benchmark_tps(model, input_ids, gen_seq_len, iteration)
    elif m in ("gemv", "gemm"):
        speed = benchmark_gemv_gemm(model, input_ids, gen_seq_len, iteration, m)
    else:
        text = tokenizer.decode(input_ids[0]) if tokenizer is not None else None
        times = []
        for _ in range(iteration):
            cleanup()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ids = input_ids
            if tokenizer is not None:
                ids = tokenizer(text, return_tensors="pt", truncation=True, max_length=input_seq_len).input_ids.to(device)
                ids = ids.expand(batch_size, -1)
            nxt = forward_prompt(model, ids).cpu()
            if tokenizer is not None:
                _ = tokenizer.decode(nxt[:1])
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1000.0)
        speed = float(np.median(times))
    key = f"{batch_size}.{input_seq_len}.{gen_seq_len}"
    data[m][key] = speed
    if get_peak_memory:
        data["peak_memory"][key] = torch.cuda.max_memory_allocated(device=device) / 1024 ** 3
        torch.cuda.reset_peak_memory_stats(device=device)
    cleanup()
    return data
