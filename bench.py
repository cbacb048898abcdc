#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native AMQ quantized-linear hot path.

Metric (BASELINE.json): batch-1 decode tok/s, Llama-2-7B AMQ mixed 2/3/4-bit (avg 3.0 bits);
dequant-GEMV achieved HBM GB/s vs peak.  A "step" is one decoded token = one pass of the hot
path (224 quantized linears + glue + fp16 lm_head) over one batch of synthetic input.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

* value      device-timed tok/s, inputs resident in HBM (CUDA-graph replay, token fed back on device)
* e2e        the same through the host-facing API: token ids come from pinned host memory each step
             and the next ids are read back to the host each step (copies inside the timed region)
* roofline   the decode GEMV family: algorithmic bytes of one step's 224 launches / their CUDA-event
             time, measured live on the launching stream, against MEASURED_PEAKS.json hbm_gbs
* cpu_baseline / --impl reference: the oracle port of the reference's torch dequant+matmul
             (GPTQLinear.forward, kernel_switch_threshold=0) on the host cores, on a bounded sample.
N > 1: the 7B path does not shard ("replicas only"): every rank decodes its own
batch-1 stream, no data-path collective, scaling "weak".  The only config that shards (config 5:
Llama-2-70B, tensor parallel over the N ranks the driver launched) is measured in the SAME run and
reported as the `tp70b` object of the line (tok/s, ms/step, launches/step, fraction of the per-GPU HBM
roofline, all-reduce kind, strong-scaling efficiency against the tp = 1 number taken in this run,
what NCCL logged about the communicator); `--workload llama70b-tp` runs config 5 alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# stdout carries EXACTLY one JSON line (the driver parses it): file descriptor 1 is pointed at stderr for the whole run, so
# that nothing a library prints there (NCCL's version banner goes to stdout even with NCCL_DEBUG_FILE set) can precede it,
# and the line itself is written to the saved descriptor.
_REAL_STDOUT = None


def _guard_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        mhz = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(mhz)}


# ------------------------------------------------------------------------------------------ CPU arm
_CPU_CACHE = {}


def cpu_reference_sample(arch, shape, threads: int, tensors=None):
    """The reference's own CPU path on a bounded sample of this workload: ONE q_proj-shaped
    4096x4096 3-bit group-128 linear, batch 1 (BASELINE.json configs[0]) through the oracle port of
    GPTQLinear.forward's torch branch (autogptq.py:245-283: shift-unpack -> fp16 scales*q - zeros ->
    matmul).  The torch path's cost is proportional to N*K whatever the bit width, so a token is
    extrapolated as t_sample * sum(N*K over the model's quantized linears) / 4096^2 (attention, norms
    and lm_head not counted, which favours the CPU arm).  Returns (seconds per token, description)."""
    import numpy as np
    import torch
    from oracle import amq_oracle as O
    torch.set_num_threads(threads)
    G, N, K, bits = 128, 4096, 4096, 3
    if tensors is not None:
        _CPU_CACHE["layer"] = tensors                 # the very tensors the GPU arm streams (run_ours)
    if "layer" not in _CPU_CACHE:
        # same recipe as the GPU arm's synthetic layers (amq_b200/model.py synthetic_native): uniform random codes,
        # realistic fp16 scale range for 3 bits, fractional zero, packed into the reference's GPTQ layout
        rs = np.random.RandomState(0)
        codes = rs.randint(0, 2 ** bits, size=(N, K)).astype(np.int64)
        qweight = O.gptq_pack_codes(codes, bits)
        scale = torch.from_numpy(rs.uniform(0.010, 0.023, size=(K // G, N)).astype(np.float32)).half()
        zero = torch.from_numpy(rs.uniform(0.5, 2 ** bits - 1.5, size=(K // G, N)).astype(np.float32)).half()
        _CPU_CACHE["layer"] = (torch.randn(1, K).half(), qweight, scale.float(), (zero * scale).float())
    x, q, sc, z = _CPU_CACHE["layer"]
    t0 = time.perf_counter()
    O.gptq_forward_torch(x, q, sc, z, bits, G)
    dt = time.perf_counter() - t0
    total_nk = sum(n * k for (n, k) in shape.linear_shape.values()) * shape.n_block
    return dt * total_nk / (N * K), ("one 4096x4096 3-bit g128 linear, batch 1 (configs[0]), torch dequant+matmul; token time "
                                    "extrapolated by N*K over the 224 quantized linears")


def cpu_reference_block(arch, shape, threads: int):
    """The reference's own CPU path on a bounded sample of this workload: ONE decoder block (block 0 of the arch: its
    seven quantized linears with their own bit widths and the shapes of amq/configs/llama.json), batch 1, through the
    oracle port of GPTQLinear.forward's torch branch (autogptq.py:245-283: shift-unpack -> fp16 scales*q - zeros ->
    matmul).  A token is n_block such blocks (attention, norms and lm_head are not counted, which favours the CPU arm).
    Returns (seconds for the block, description)."""
    import numpy as np
    import torch
    from oracle import amq_oracle as O
    from amq_b200.arch import LINEARS
    from amq_b200.model import _SCALE_RANGE
    torch.set_num_threads(threads)
    G = 128
    if "block" not in _CPU_CACHE:
        # same recipe as the GPU arm's synthetic layers (amq_b200/model.py synthetic_native): uniform random codes, the
        # realistic fp16 scale range of the bit width, fractional zero, packed into the reference's GPTQ layout
        rs = np.random.RandomState(0)
        layers = []
        for name in LINEARS:
            bits = int(arch[name][0])
            N, K = shape.linear_shape[name]
            qweight = O.gptq_pack_codes(rs.randint(0, 2 ** bits, size=(N, K)).astype(np.int64), bits)
            lo, hi = _SCALE_RANGE[bits]
            scale = torch.from_numpy(rs.uniform(lo, hi, size=(K // G, N)).astype(np.float32)).half()
            zero = torch.from_numpy(rs.uniform(0.5, 2 ** bits - 1.5, size=(K // G, N)).astype(np.float32)).half()
            layers.append((name, bits, torch.randn(1, K).half(), qweight, scale.float(), (zero * scale).float()))
        _CPU_CACHE["block"] = layers
    layers = _CPU_CACHE["block"]
    t0 = time.perf_counter()
    for _, bits, x, q, sc, z in layers:
        O.gptq_forward_torch(x, q, sc, z, bits, G)
    dt = time.perf_counter() - t0
    desc = ("one decoder block of the workload (" + ", ".join(f"{n.split('.')[-1]} {b}b" for n, b, *_ in layers) +
            f"; 1/{shape.n_block} of a token), batch 1, torch dequant+matmul")
    return dt, desc


def run_reference(args, shape, arch):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    times = []
    for i in range(args.warmup + args.steps):
        t, desc = cpu_reference_block(arch, shape, threads)
        if i >= args.warmup:
            times.append(t)
    step_time = sum(times) / len(times)            # one block = 1 / n_block of a token
    tok_s = 1.0 / (step_time * shape.n_block)
    line = {
        "impl": "reference", "metric": "batch-1 decode tok/s, Llama-2 7B AMQ 3-bit avg", "value": tok_s, "unit": "tok/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_time * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"{shape.name} random-init, AMQ mixed 2/3/4-bit avg 3.0 (synthetic arch, seed 0), batch-1 decode",
                   "timing": f"host wall clock; each step = ONE of the {shape.n_block} decoder blocks (its 7 quantized linears) on the "
                             f"CPU = 1/{shape.n_block} token; value = 1 / ({shape.n_block} x step time), extrapolated over the identical blocks",
                   "tokens_per_step": 1.0 / shape.n_block},
        "cpu_baseline": {"value": tok_s, "unit": "tok/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": tok_s, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


# ------------------------------------------------------------------------------------------ GPU arm
def cpu_sample_from_model(model):
    """A 4096 x 4096 3-bit linear of the GPU arm's own model (first q/k/v/o projection drawn at 3 bits), brought to the
    reference's GPTQ layout for the CPU baseline: codes unpacked from the native buffer, fp16 scale / zero*scale read
    from its records."""
    import numpy as np
    import torch
    from amq_b200 import _lib, ops
    from oracle import amq_oracle as O
    for L in model.layers:
        for name in ("self_attn.q_proj", "self_attn.o_proj", "self_attn.k_proj", "self_attn.v_proj"):
            bits, nat, N, K = L[name]
            if bits == 3 and N == 4096 and K == 4096:
                codes = ops.unpack_codes(nat, bits, _lib.LAYOUT_NATIVE, N, K, 128).cpu().numpy().astype(np.int64)
                rec = bits * 512 + 128
                meta = nat.reshape(-1, rec)[:, bits * 512:].contiguous().view(torch.float16).reshape(N // 32, K // 128, 32, 2)
                scales = meta[..., 0].permute(1, 0, 2).reshape(K // 128, N).float().cpu()
                zeros = meta[..., 1].permute(1, 0, 2).reshape(K // 128, N).float().cpu()
                x = model.embed[1:2].float().cpu().half() * 50.0      # an activation row of the model's own embedding scale
                return (x, O.gptq_pack_codes(codes, bits), scales, zeros)
    return None


def gemv_roofline(model, iters: int = 5):
    """Time one step's worth of decode-GEMV launches alone (same problems, same order, PDL on, CUDA
    graph, events on the launching stream).  The 2.4 GB of packed weights exceed L2 (126 MB)."""
    import torch
    s = torch.cuda.Stream(device=model.dev)
    n_launch = 0
    with torch.cuda.stream(s):
        def launches():
            n = 0
            for P in model._plan:
                for key, cnt in (("qkv", 3), ("o", 1), ("gu", 2), ("down", 1)):
                    model._gemv(P[key], cnt)
                    n += len({P[key][i].prologue for i in range(cnt)})
            return n
        n_launch = launches()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            launches()
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(iters):
            g.replay()
        e1.record(s)
        s.synchronize()
    ms = e0.elapsed_time(e1) / iters
    alg = model.algorithmic_bytes_per_token()["linears"] + 7 * model.n_block * 2 * 2 * model.H   # + x / y traffic (approx.)
    return ms, n_launch, alg


def run_ours(args, shape, arch):
    import torch
    import torch.distributed as dist
    from amq_b200.arch import get_bits_usage
    from amq_b200.model import QuantDecoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nccl_log = None
    if world > 1:
        # leave NCCL's own logging on (to a file: stdout carries exactly one JSON line) so that the communicator size
        # and transport it reports can be quoted next to the tensor-parallel numbers
        nccl_log = f"/tmp/amqb_nccl_{os.getpid()}_%h_%p.log"
        os.environ["NCCL_DEBUG"] = "INFO"                 # the image presets NCCL_DEBUG=VERSION: override, the output goes to the file
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,GRAPH,ENV")
        os.environ.setdefault("NCCL_DEBUG_FILE", nccl_log)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    B = 1
    model = QuantDecoder(shape, arch, batch=B, max_seq=max(512, args.warmup + 2 * args.steps + 8), device=f"cuda:{local}", seed=rank)
    model.capture()
    launches_per_step, model_pdl = model.launches_per_step, model.pdl

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K graph replays, tokens fed back on the device
    model.reset()
    model.tokens.fill_(1)
    for _ in range(args.warmup):
        model.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof = os.environ.get("AMQB_PROFILE") == "1"        # ncu --profile-from-start off: only the timed steps
    if prof:
        torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    for _ in range(args.steps):
        model.step()
    e1.record()
    barrier()
    if prof:
        torch.cuda.cudart().cudaProfilerStop()
    ms_dev = e0.elapsed_time(e1)

    # ---- end to end through the host-facing call: pinned host ids in, ids out, every step
    host_in = torch.ones(B, dtype=torch.int64).pin_memory()
    host_out = torch.zeros(B, dtype=torch.int64).pin_memory()
    model.reset()                                   # same positions (KV lengths) as the device-timed steps above
    for _ in range(args.warmup):
        model.step_host(host_in, host_out)
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        # H2D of this step's input ids, the decode step, D2H of the generated ids: one captured graph, then a
        # synchronise (QuantDecoder.step_host); the host owns the loop and feeds the ids back
        model.step_host(host_in, host_out)
        host_in.copy_(host_out)
    e3.record()
    barrier()
    ms_e2e = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peak, peak_src = _peaks()
        g_ms, g_launch, g_bytes = gemv_roofline(model)
        achieved = g_bytes / (g_ms * 1e-3) / 1e9
        bytes_tok = model.algorithmic_bytes_per_token()
        cpu_tensors = cpu_sample_from_model(model)
    del model
    torch.cuda.empty_cache()

    # ---- config 5 in the same run: Llama-2-70B tensor-parallel over the ranks the driver launched (N = 1: one GPU),
    # plus, at N > 1, the tp = 1 number on rank 0 so that the strong-scaling efficiency is self-contained
    from amq_b200 import tp as tpmod
    tp_steps, tp_warm = max(8, min(args.steps, 48)), max(3, min(args.warmup, 6))
    tp_rec = tp1_rec = None
    if os.environ.get("AMQB_SKIP_TP70B") != "1":
        tp_rec = tpmod.measure_tp70b(tp_steps, tp_warm, world, rank, local)
        if world > 1:
            tp1_rec = tpmod.measure_tp70b(max(8, tp_steps // 2), tp_warm, world, rank, local, tp=1)
    if rank == 0 and tp_rec is not None:
        base = tp1_rec if tp1_rec is not None else tp_rec
        tp_rec["tp1_tok_s"] = base["tok_s"]
        tp_rec["tp1_prompt_pass_ms"] = base["prompt_pass_ms"]
        tp_rec["strong_scaling_efficiency"] = tp_rec["tok_s"] / (world * base["tok_s"])
        if nccl_log is not None:
            tp_rec["nccl"] = tpmod.nccl_log_summary(f"/tmp/amqb_nccl_{os.getpid()}_*.log")

    if rank == 0:
        cpu_runs = [cpu_reference_sample(arch, shape, os.cpu_count() or 1, cpu_tensors if i == 0 else None) for i in range(12)]
        cpu_desc = cpu_runs[0][1]
        cpu_tok_s = 1.0 / (sum(r[0] for r in cpu_runs[2:]) / len(cpu_runs[2:]))
        tok_s = world * B * args.steps / (ms_dev * 1e-3)
        # dram bytes per launch (average over one layer's four launches): NOT measured in this run — read from the
        # committed `ncu --set full` capture of the same launches (newest round first)
        traffic, traffic_src = None, None
        for name in ("r02_ncu_gemv_traffic.json", "r01_ncu_gemv_traffic.json"):
            tp_ = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tp_):
                with open(tp_) as f:
                    traffic = json.load(f).get("traffic_bytes_per_launch_avg")
                traffic_src = f"profiles/{name} (ncu --set full capture of one layer's launches; not measured in this run)"
                break
        line = {
            "metric": "batch-1 decode tok/s, Llama-2 7B AMQ 3-bit avg", "value": tok_s, "unit": "tok/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": f"{shape.name} random-init, AMQ mixed 2/3/4-bit, bits_usage "
                                   f"{get_bits_usage({'linear': arch}, shape.config()):.2f} (synthetic arch drawn like the search space, "
                                   "seed 0; func.py accounting incl. 0.25 b scale/zero), batch-1 decode, group 128",
                       "l2": "inputs larger than L2: every step streams %.2f GB of packed weights + fp16 lm_head" % (bytes_tok["total"] / 1e9),
                       "parallelism": "replicas only (no data-path collective)" if world > 1 else "single GPU",
                       "launches_per_step": launches_per_step, "cuda_graph": True, "pdl": model_pdl},
            "clocks": clocks,
            "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "tok/s",
                    "h2d_bytes_per_step": 8 * B, "d2h_bytes_per_step": 8 * B},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": g_bytes / g_launch, "peak_source": peak_src,
                         "kernel": "gemv_mma_kernel<MB,kind,prologue> (IMMA decode GEMV family, %d launches per step)" % g_launch,
                         "algorithmic_bytes_per_step": g_bytes, "avg_launch_us": g_ms * 1e3 / g_launch,
                         "step_frac_of_weight_roofline": (bytes_tok["total"] / (peak * 1e9)) / (ms_dev / args.steps * 1e-3)},
            "cpu_baseline": {"value": cpu_tok_s, "unit": "tok/s", "cores": os.cpu_count(), "kind": "port", "sample": cpu_desc},
            "tp70b": tp_rec,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="llama7b", choices=["llama7b", "llama70b-tp"])
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    from amq_b200.arch import MODELS, sample_arch
    if args.workload == "llama70b-tp":
        from amq_b200 import tp
        return tp.bench_main(args)
    shape = MODELS["Llama-2-7b-hf"]
    arch = sample_arch(shape, 3.0, seed=0)
    if args.impl == "reference":
        return run_reference(args, shape, arch)
    return run_ours(args, shape, arch)


if __name__ == "__main__":
    main()
