"""The oracle against the committed golden fixtures (made from the unmodified
reference by oracle/gen_golden.py).  CPU only."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import amq_oracle as O

G = 128


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "bitpack_*.npz"))))
def test_bitpack_roundtrip(path):
    d = np.load(path)
    nbits = int(os.path.basename(path).split("_")[1][0])
    codes, packed = d["codes"], d["packed"]
    assert np.array_equal(O.hqq_pack(codes, nbits), packed)
    assert np.array_equal(O.hqq_unpack(packed, nbits, rows=codes.shape[0]), codes)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "linear_*.npz"))))
def test_linear_fixture(path):
    d = np.load(path)
    nbits = int(os.path.basename(path).split("_")[1][0])
    W = torch.from_numpy(d["W"])
    N, K = W.shape
    codes, scale, zero, n_it = O.hqq_quantize(W, nbits, G)
    assert np.array_equal(codes, d["codes"])
    assert np.array_equal(scale.numpy(), d["hqq_scale"]) and np.array_equal(zero.numpy(), d["hqq_zero"])
    assert n_it == int(d["solver_iters"])
    assert np.array_equal(O.hqq_pack(codes, nbits), d["hqq_Wq"])
    s16, z16 = scale.half(), zero.half()
    W_deq = O.hqq_dequantize(codes, s16, z16, (N, K))
    assert np.array_equal(W_deq.numpy(), d["W_deq"])
    q = O.gptq_codes_from_weight(W_deq, s16.reshape(N, -1), z16.reshape(N, -1), G)
    assert np.array_equal(q.astype(np.uint8), codes.reshape(N, K))
    assert np.array_equal(O.gptq_pack_codes(q, nbits), d["gptq_qweight"])
    assert np.array_equal(O.gptq_unpack(d["gptq_qweight"], nbits), q.T.astype(np.uint8))
    assert np.array_equal(O.gptq_unpack_fast(d["gptq_qweight"], nbits), q.T.astype(np.uint8))
    scales = torch.from_numpy(d["gptq_scales"])
    zeros = torch.from_numpy(d["gptq_zeros"])
    for M in (1, 5):
        x = torch.from_numpy(d[f"x{M}"])
        y = O.gptq_forward_torch(x, d["gptq_qweight"], scales, zeros, nbits, G)
        assert np.array_equal(y.numpy(), d[f"y{M}_ref_fp16"])
        y32 = O.gptq_forward_fp32(x, d["gptq_qweight"], scales, zeros, nbits, G)
        np.testing.assert_allclose(y32.numpy(), d[f"y{M}_fp32"], rtol=1e-6, atol=1e-7)
        # the reference's own fp16 path sits inside the 1e-3 max-rel band
        assert O.max_rel(y, y32) < 1e-3
    if nbits == 4:
        assert np.array_equal(O.ft_pack_intweight(q), d["ft_qweight"])
        assert np.array_equal(O.ft_unpack(d["ft_qweight"]), q.astype(np.uint8))
        x = torch.from_numpy(d["x5"])
        y_ft = O.ft_forward_fp32(x, d["ft_qweight"], scales.half().float(), -zeros.half().float(), G)
        np.testing.assert_allclose(y_ft.numpy(), d["y5_fp32"], rtol=1e-5, atol=1e-6)


def test_arch_selection(golden_dir):
    with open(os.path.join(golden_dir, "arch_stats.json")) as f:
        d = json.load(f)
    assert O.select_arch(d["stats"], d["target_bits"]) == d["expected"]


@pytest.mark.parametrize("bits", [2, 3, 4])
def test_oracle_roundtrip_properties(bits):
    """Size-independent properties of the restated layouts on seeded random codes (beyond the committed fixtures):
    every packer is inverted exactly by its unpacker, including ragged 3-bit row counts (padding to a multiple of 10,
    bitpack.py:69-91), the smallest legal shapes, and the two 3-bit codes of every 32 that straddle a word boundary
    (autogptq.py:133-151)."""
    rs = np.random.RandomState(100 + bits)
    for R in ([1, 9, 10, 13, 130, 257] if bits == 3 else [4, 8, 64, 256]):
        codes = rs.randint(0, 2 ** bits, size=(R, 128)).astype(np.uint8)
        packed = O.hqq_pack(codes, bits)
        assert packed.shape[0] == ((R + 9) // 10 if bits == 3 else R // (8 // bits))
        assert np.array_equal(O.hqq_unpack(packed, bits, rows=R), codes)
    for N, K in [(2, 32), (6, 96), (64, 256)]:
        codes = rs.randint(0, 2 ** bits, size=(N, K)).astype(np.uint8)
        qw = O.gptq_pack_codes(codes, bits)
        assert qw.shape == (K * bits // 32, N) and qw.dtype == np.int32
        assert np.array_equal(O.gptq_unpack(qw, bits), codes.T)
        assert np.array_equal(O.gptq_unpack_fast(qw, bits), codes.T)
    # extreme codes: all-ones fields must not bleed into their neighbours
    full = np.full((4, 64), 2 ** bits - 1, dtype=np.uint8)
    full[:, ::2] = 0
    assert np.array_equal(O.gptq_unpack(O.gptq_pack_codes(full, bits), bits), full.T)
    if bits == 4:
        for N, K in [(4, 64), (8, 192), (64, 256)]:
            codes = rs.randint(0, 16, size=(N, K)).astype(np.uint8)
            q = O.ft_pack_intweight(codes)
            assert q.shape == (N // 4, K) and q.dtype == np.int16
            assert np.array_equal(O.ft_unpack(q), codes)


def test_oracle_forward_linearity_and_bits_accounting():
    """The restated torch forward is linear in x (fp32 form) and agrees with a dense matmul of the dequantised weight;
    get_bits_usage reproduces func.py:101-114 on a hand-computed case."""
    rs = np.random.RandomState(7)
    bits, N, K, G = 3, 32, 256, 128
    codes = rs.randint(0, 8, size=(N, K)).astype(np.uint8)
    qw = O.gptq_pack_codes(codes, bits)
    scales = torch.from_numpy(rs.uniform(0.01, 0.02, size=(K // G, N)).astype(np.float32)).half().float()
    zeros = torch.from_numpy(rs.uniform(0.02, 0.1, size=(K // G, N)).astype(np.float32)).half().float()
    x1, x2 = torch.randn(1, K).half(), torch.randn(1, K).half()
    y1 = O.gptq_forward_fp32(x1, qw, scales, zeros, bits, G)
    y2 = O.gptq_forward_fp32(x2, qw, scales, zeros, bits, G)
    y12 = O.gptq_forward_fp32((x1.float() + x2.float()), qw, scales, zeros, bits, G)
    assert torch.allclose(y1 + y2, y12, rtol=1e-5, atol=1e-5)
    W = (torch.from_numpy(codes.T.astype(np.float32)).reshape(K // G, G, N) * scales[:, None, :] - zeros[:, None, :]).reshape(K, N)
    assert torch.allclose(y1, x1.float() @ W, rtol=1e-5, atol=1e-5)
    cfg = {"linear_shape": {"a": [4, 256], "b": [8, 128]}, "model_numel": 4 * 256 + 8 * 128}
    arch = {"linear": {"a": [2, 4], "b": [3, 3]}}
    want = (4 * 256 * (2.25 + 4.25) + 8 * 128 * (3.25 + 3.25)) / (4 * 256 + 8 * 128)
    assert abs(O.get_bits_usage(arch, cfg) - want) < 1e-12


def test_sleef_powf_restatement_matches_torch(tmp_path):
    """oracle/sleef_powf.c restates Sleef's powf (the kernel behind torch's CPU x.pow(p - 1) in the HQQ solver's shrink
    operator, optimize.py:96-108) in scalar C; the CUDA solver carries the same operation sequence.  Pin: bit-equal to
    torch.pow on 2 M positive inputs spanning the solver's range (single thread, numel % 16 == 0: torch's vector path)."""
    import ctypes
    import subprocess
    import torch
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "sleef_powf.c")
    so = tmp_path / "libsleefpow.so"
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), src, "-lm"], check=True)
    L = ctypes.CDLL(str(so))
    torch.manual_seed(0)
    a = torch.cat([torch.rand(1_000_000) * 0.05 + 1e-7, torch.exp(torch.empty(1_000_000).uniform_(-30.0, 2.0))]).float()
    y = 0.7 - 1
    nt = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        p = a.pow(y).numpy()
    finally:
        torch.set_num_threads(nt)
    x = a.numpy().copy()
    out = np.empty_like(x)
    L.sleef_powf(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_float(np.float32(y)), out.ctypes.data_as(ctypes.c_void_p),
                 ctypes.c_long(x.size))
    assert np.array_equal(p.view(np.int32), out.view(np.int32))
