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
