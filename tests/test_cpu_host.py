"""CPU-only checks: the C-ABI library loads and exports every symbol include/amqb.h declares,
host-side logic (arch JSON handling, module constructors / state-dict contract, sharding plan),
loud failure without CUDA, and the world-size-2 rendezvous path over gloo."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "amqb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(amqb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from amq_b200.build import build
    path = build()
    lib = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.amqb_last_error_string.restype = ctypes.c_char_p
    assert lib.amqb_version() >= 100
    # pure host-side entry points may be called without a GPU
    lib.amqb_native_bytes.restype = ctypes.c_size_t
    assert lib.amqb_native_bytes(3, 4096, 4096) == 128 * 32 * 1664
    assert lib.amqb_native_bytes(2, 4096, 4096) * 8 == 4096 * 4096 * 9 // 4      # 2.25 bits / weight
    assert lib.amqb_native_bytes(3, 100, 4096) == 0 and lib.amqb_native_bytes(5, 64, 128) == 0
    # exchange buffers of the tensor-parallel all-reduces: header + one flag line per (rank, CTA) + two generations of slots
    lib.amqb_ar_rows_buffer_bytes.restype = ctypes.c_size_t
    lib.amqb_ar_buffer_bytes.restype = ctypes.c_size_t
    assert lib.amqb_ar_rows_buffer_bytes(63 * 8192, 8) == 1024 + 128 * 8 * 64 + 2 * 8 * 63 * 8192 * 2
    assert lib.amqb_ar_rows_buffer_bytes(0, 8) == 0 and lib.amqb_ar_rows_buffer_bytes(8, 17) == 0
    assert lib.amqb_ar_buffer_bytes(8192, 8) > 2 * 8 * 8192 * (2 + 8)


def test_native_layout_index_math(tmp_path):
    """Host emulation of the byte-field layout (layout.cuh): pack -> masked IMMA registers x integer activation
    slots == 2^smax * sum(code * X) for every row and bit width; every code bit stored exactly once."""
    exe = str(tmp_path / "layout_check")
    subprocess.run(["g++", "-O1", "-std=c++17", os.path.join(ROOT, "tests", "native", "layout_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "layout ok" in out.stdout, out.stdout + out.stderr


def test_no_cpu_fallback():
    import amq_b200
    from amq_b200 import ops
    layer = amq_b200.GPTQLinear(3, 128, 256, 64, bias=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.zeros(1, 256, dtype=torch.float16))
    with pytest.raises(RuntimeError):
        ops.gemv(3, torch.zeros(16, dtype=torch.uint8), torch.zeros(1, 128, dtype=torch.float16), 32, 128)
    with pytest.raises(RuntimeError):
        amq_b200.Quantizer.quantize(torch.zeros(64, 128), nbits=3, group_size=128, axis=1, device="cpu")


def test_module_contract_matches_reference_buffers():
    """Buffer names / shapes / dtypes of the drop-in modules (autogptq.py:55-82, ft.py:76-94)."""
    import amq_b200
    for bits in (2, 3, 4):
        m = amq_b200.GPTQLinear(bits, 128, 4096, 11008, bias=True)
        sd = m.state_dict()
        assert set(sd) == {"qweight", "zeros", "scales", "bias"}
        assert sd["qweight"].shape == (4096 // 32 * bits, 11008) and sd["qweight"].dtype == torch.int32
        assert sd["scales"].shape == (32, 11008) and sd["scales"].dtype == torch.float32
        assert sd["zeros"].shape == (32, 11008) and sd["bias"].dtype == torch.float16
        assert m.maxq == 2 ** bits - 1 and m.half_indim == 2048 and m.QUANT_TYPE == "cuda-old"
    with pytest.raises(NotImplementedError):
        amq_b200.GPTQLinear(5, 128, 256, 256, False)
    f = amq_b200.FT_QuantLinear(4, 4096, 4096, False, torch.float16, 128, "q_proj")
    sd = f.state_dict()
    assert set(sd) == {"qweight", "scales", "scaled_zeros"}
    assert sd["qweight"].shape == (1024, 4096) and sd["qweight"].dtype == torch.int16
    assert sd["scales"].shape == (32, 4096) and sd["scales"].dtype == torch.float16
    with pytest.raises(AssertionError):
        amq_b200.FT_QuantLinear(3, 4096, 4096, False, torch.float16, 128, "x")
    # a reference-format state dict loads and invalidates the native repack
    m = amq_b200.GPTQLinear(3, 128, 256, 64, bias=False)
    m.load_state_dict({"qweight": torch.ones(24, 64, dtype=torch.int32), "zeros": torch.zeros(2, 64), "scales": torch.ones(2, 64)})
    assert m._native_ok is None and int(m.qweight[0, 0]) == 1
    cfg = amq_b200.BaseQuantizeConfig(nbits=4, group_size=128)
    assert cfg["weight_quant_params"]["round_zero"] is True and cfg["weight_quant_params"]["axis"] == 1
    assert amq_b200.BaseQuantizeConfig(nbits=3, group_size=128)["weight_quant_params"]["round_zero"] is False


def test_arch_handling(tmp_path, golden_dir):
    from amq_b200 import arch
    with open(os.path.join(golden_dir, "arch_stats.json")) as f:
        d = json.load(f)
    assert arch.select_arch(d["stats"], d["target_bits"]) == d["expected"]
    shape = arch.MODELS["Llama-2-7b-hf"]
    a = arch.sample_arch(shape, 3.0, seed=0)
    assert set(a) == set(arch.LINEARS) and all(len(v) == 32 for v in a.values())
    assert abs(arch.get_bits_usage({"linear": a}, shape.config()) - 3.0) < 0.05
    # uniform 3-bit = 3.25 bits/weight share of the linears (func.py:101-114)
    u = arch.load_arch(None, 3, 32)
    lin_numel = sum(n * k for n, k in shape.linear_shape.values()) * 32
    assert abs(arch.get_bits_usage({"linear": u}, shape.config()) - 3.25 * lin_numel / shape.model_numel) < 1e-9
    p = tmp_path / "iter_1.stats"
    arch.make_stats_file(str(p), shape, 3.0, n=4)
    sel = arch.load_arch(str(p), 3.0, 32)
    assert abs(arch.get_bits_usage({"linear": sel}, shape.config()) - 3.0) < 0.05
    with pytest.raises(FileNotFoundError):
        arch.load_arch(str(tmp_path / "missing"), 3.0, 32)
    # the shape tables are the reference's (amq/configs/*.json)
    assert arch.MODELS["Llama-2-70b-hf"].linear_shape["self_attn.k_proj"] == [1024, 8192]
    assert arch.MODELS["Mistral-7B-v0.3"].linear_shape["mlp.down_proj"] == [4096, 14336]
    assert arch.MODELS["Qwen2.5-7B"].linear_shape["self_attn.v_proj"] == [512, 3584]


def test_shard_plan_and_buffer_slicing():
    from amq_b200 import tp
    from amq_b200.arch import MODELS
    from oracle import amq_oracle as O
    shape = MODELS["Llama-2-70b-hf"]
    for world in (1, 2, 4, 8):
        plan = tp.shard_plan(shape, world)
        assert plan["self_attn.k_proj"]["N"] == 1024 // world and plan["mlp.down_proj"]["K"] == 28672 // world
    with pytest.raises(ValueError):
        tp.shard_plan(shape, 16)          # 8 kv heads do not split 16 ways
    # slicing reference-layout buffers commutes with unpacking (column and row split, 3-bit)
    rs = np.random.RandomState(0)
    N, K, bits = 64, 512, 3
    codes = rs.randint(0, 8, size=(N, K))
    qw = torch.from_numpy(O.gptq_pack_codes(codes, bits))
    sc = torch.rand(K // 128, N)
    ze = torch.rand(K // 128, N)
    for rank in range(2):
        q, s, z = tp.shard_gptq_buffers(qw, sc, ze, bits, "column", rank, 2)
        assert np.array_equal(O.gptq_unpack(q.numpy(), bits), codes[rank * 32:(rank + 1) * 32].T)
        q, s, z = tp.shard_gptq_buffers(qw, sc, ze, bits, "row", rank, 2)
        assert np.array_equal(O.gptq_unpack(q.numpy(), bits), codes[:, rank * 256:(rank + 1) * 256].T)
        assert torch.equal(s, sc[rank * 2:(rank + 1) * 2])


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from amq_b200 import tp
from oracle import amq_oracle as O
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
hs = tp.exchange_handles(bytes([rank]) * 64)
assert [h[0] for h in hs] == list(range(world)) and all(len(h) == 64 for h in hs)
# row-parallel linear: every rank multiplies its K slice, the all-reduce restores the full product
rs = np.random.RandomState(0)
N, K, bits = 64, 512, 4
codes = rs.randint(0, 16, size=(N, K))
qw = torch.from_numpy(O.gptq_pack_codes(codes, bits))
sc = torch.from_numpy(rs.uniform(0.01, 0.02, size=(K // 128, N)).astype(np.float32))
ze = torch.from_numpy(rs.uniform(0.02, 0.1, size=(K // 128, N)).astype(np.float32))
x = torch.from_numpy(rs.randn(1, K).astype(np.float32)).half()
q, s, z = tp.shard_gptq_buffers(qw, sc, ze, bits, "row", rank, world)
k0, k1 = K * rank // world, K * (rank + 1) // world
part = O.gptq_forward_fp32(x[:, k0:k1], q.numpy(), s, z, bits, 128)
dist.all_reduce(part)
full = O.gptq_forward_fp32(x, qw.numpy(), sc, ze, bits, 128)
assert torch.allclose(part, full, rtol=1e-5, atol=1e-5)
# prompt pass of a sharded MLP block over M rows: column-parallel gate / up (this rank's columns), row-parallel down
# (its k groups), ONE all-reduce of the [M, hidden] partial sums (what amqb_allreduce_rows_f16 does over peer memory)
M, H, I = 5, 128, 512
def lin(n, k, seed):
    r = np.random.RandomState(seed)
    c = r.randint(0, 8, size=(n, k))
    return (torch.from_numpy(O.gptq_pack_codes(c, 3)), torch.from_numpy(r.uniform(0.01, 0.02, size=(k // 128, n)).astype(np.float32)),
            torch.from_numpy(r.uniform(0.02, 0.1, size=(k // 128, n)).astype(np.float32)))
gate, up, down = lin(I, H, 1), lin(I, H, 2), lin(H, I, 3)
xm = torch.from_numpy(np.random.RandomState(4).randn(M, H).astype(np.float32)).half()
fwd = lambda xx, w: O.gptq_forward_fp32(xx, w[0].numpy(), w[1], w[2], 3, 128)
act = lambda g, u: (torch.nn.functional.silu(g) * u).half()
want = fwd(act(fwd(xm, gate), fwd(xm, up)), down)
plan = tp.shard_plan(tp.ModelShape("tiny", H, I, 4, 4, 1, 64, head_dim=64), world)
assert plan["mlp.gate_proj"] == {"N": I // world, "K": H, "split": "column"} and plan["mlp.down_proj"]["split"] == "row"
g_loc = tp.shard_gptq_buffers(*gate, 3, "column", rank, world)
u_loc = tp.shard_gptq_buffers(*up, 3, "column", rank, world)
d_loc = tp.shard_gptq_buffers(*down, 3, "row", rank, world)
part = fwd(act(fwd(xm, g_loc), fwd(xm, u_loc)), d_loc)
assert part.shape == (M, H)
dist.all_reduce(part)
assert torch.allclose(part, want, rtol=1e-4, atol=1e-4), float((part - want).abs().max())
dist.destroy_process_group()
print("ok", rank)
'''


def test_gloo_world2_row_parallel(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_speed_benchmark_cli_mirrors_reference_flags():
    """amq_speed_benchmark.py keeps the reference's command line (amq/amq_speed_benchmark.py:103-125: names, types,
    defaults) and refuses to run without a GPU; benchmark_speed keeps the reference's signature (speed.py:130)."""
    import inspect
    sys.path.insert(0, ROOT)
    import amq_speed_benchmark as cli
    from amq_b200.utils import speed
    p = cli.build_parser()
    a = p.parse_args([])
    ref_defaults = {"model_path": "meta-llama", "model_name": "Llama-2-7b-hf", "save_path": "/SSD/hqq", "use_ft": False,
                    "batch_size": 1, "seq_length": 64, "gen_length": 128, "tps": False, "gemm": False, "gemv": False,
                    "ttft": False, "memory": False, "peak_memory": False, "target_bits": 4, "arch_path": None,
                    "file_name": None}
    for k, v in ref_defaults.items():
        assert getattr(a, k) == v, k
    a = p.parse_args(["--tps", "--gemv", "--target_bits", "3", "--arch_path", "x.stats", "--batch_size", "4"])
    assert a.tps and a.gemv and a.target_bits == 3.0 and a.arch_path == "x.stats" and a.batch_size == 4
    sig = inspect.signature(speed.benchmark_speed)
    assert list(sig.parameters) == ["model", "tokenizer", "use_ft", "iteration", "sizes", "mode", "get_peak_memory"]
    assert sig.parameters["sizes"].default == (1, 128, 128) and sig.parameters["mode"].default == "TPS"
    with pytest.raises(AssertionError):
        speed.benchmark_speed(None, mode="latency")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            cli.main(["--tps"])


def test_bench_reference_arm_prints_exactly_one_json_line():
    """The driver parses bench.py's stdout: exactly one JSON line, whatever libraries print (bench.py points file
    descriptor 1 at stderr for the run and writes the line to the saved descriptor).  The reference arm is the CPU arm, so
    it runs here."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tok/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
    # a step is one decoder block = 1 / 32 of a token, timed for real: the line's ms_per_step is the measured step, and
    # value is tokens per step over the step time
    assert abs(d["value"] - d["config"]["tokens_per_step"] / (d["ms_per_step"] * 1e-3)) <= 1e-9 * d["value"]
    assert d["config"]["tokens_per_step"] == 1 / 32 and "q_proj" in d["cpu_baseline"]["sample"]
