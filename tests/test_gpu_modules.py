"""GPU tests of the host-side mirror of the reference interface: Quantizer / HQQLinear (config 4),
prepare_for_inference with the gptq / ft backends and the state-dict cache files, the drop-in modules'
forward for decode and prefill row counts, against the oracle and the golden fixtures."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import amq_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
G = 128


@pytest.fixture(scope="module")
def amq():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import amq_b200
    return amq_b200


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "linear_*.npz"))))
def test_quantizer_bit_exact_against_reference_fixture(amq, path):
    """Quantizer.quantize on the GPU (fp32 solver arithmetic = the reference's CPU branch) against the reference's own
    output (golden): codes, packed W_q, scale, zero and the number of solver iterations are BIT-EXACT (the solver carries
    Sleef's powf and ATen's row-sum order op by op, csrc/hqq_quant.cu)."""
    d = np.load(path)
    bits = int(os.path.basename(path).split("_")[1][0])
    W = torch.from_numpy(d["W"]).cuda()
    N, K = W.shape
    from amq_b200 import ops
    codes, scale, zero, iters, W_q = ops.hqq_quantize(W, bits, G, solver_dtype=torch.float32, packed=True)
    assert int(iters.item()) == int(d["solver_iters"])
    assert np.array_equal(codes.cpu().numpy(), d["codes"])
    assert np.array_equal(W_q.cpu().numpy(), d["hqq_Wq"])
    assert np.array_equal(scale.cpu().numpy(), d["hqq_scale"])
    assert np.array_equal(zero.cpu().numpy(), d["hqq_zero"])
    # the unpacked-codes entry point gives the same
    c2, s2, z2, it2 = ops.hqq_quantize(W, bits, G, solver_dtype=torch.float32)
    assert torch.equal(c2, codes) and torch.equal(s2, scale) and torch.equal(z2, zero) and int(it2.item()) == int(iters.item())
    # through the reference-facing class
    amq.Quantizer.solver_dtype = torch.float32
    try:
        cfg = amq.BaseQuantizeConfig(nbits=bits, group_size=G)["weight_quant_params"]
        Wq2, meta = amq.Quantizer.quantize(W, device="cuda", compute_dtype=torch.float16, **cfg)
    finally:
        amq.Quantizer.solver_dtype = None
    assert np.array_equal(Wq2.cpu().numpy(), d["hqq_Wq"]) and str(Wq2.dtype).endswith(str(d["hqq_Wq"].dtype))
    assert np.array_equal(meta["scale"].cpu().numpy(), d["hqq_scale"]) and np.array_equal(meta["zero"].cpu().numpy(), d["hqq_zero"])
    meta16 = dict(meta, scale=meta["scale"].half(), zero=meta["zero"].half(), compute_dtype=torch.float16)
    assert np.array_equal(amq.Quantizer.dequantize(Wq2, meta16).cpu().numpy(), d["W_deq"])


@pytest.mark.parametrize("bits", [2, 3, 4])
def test_quantizer_bit_exact_against_cpu_oracle_large(amq, bits):
    """Same bit-exactness on a model-sized layer (1024 x 4096: 32768 groups, several solver blocks and reduction
    partials) against the CPU oracle run here (single-threaded so that torch's vectorised pow covers every element)."""
    from amq_b200 import ops
    torch.manual_seed(10 + bits)
    N, K = 1024, 4096
    W = (torch.randn(N, K) * 0.02).half()
    nt = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        codes, scale, zero, n_it = O.hqq_quantize(W, bits, G)
    finally:
        torch.set_num_threads(nt)
    c, s, z, it, W_q = ops.hqq_quantize(W.cuda(), bits, G, solver_dtype=torch.float32, packed=True)
    assert int(it.item()) == n_it
    assert np.array_equal(c.cpu().numpy(), codes)
    assert torch.equal(s.cpu(), scale) and torch.equal(z.cpu(), zero)
    assert np.array_equal(W_q.cpu().numpy(), O.hqq_pack(codes, bits))


@pytest.mark.parametrize("bits", [2, 3, 4])
def test_quantizer_fp16_solver_mode(amq, bits):
    """The reference's CUDA branch runs the solver in fp16 (optimize.py:231); tests/golden/hqq_fp16solver_* hold what the
    reference's own step function gives on fp16 CPU tensors.  Same iteration count; codes / zero equal except where the
    CUDA and CPU fp16 kernels round a pow or a 128-mean differently in the last bit (a flipped intermediate code moves that
    group's zero by k / 128): bounded, not bit-exact — the fp32 mode above is the pinned one."""
    from amq_b200 import ops
    d = np.load(os.path.join(GOLD, f"hqq_fp16solver_{bits}bit.npz"))
    W = torch.from_numpy(d["W"]).cuda()
    c, s, z, it, W_q = ops.hqq_quantize(W, bits, G, solver_dtype=torch.float16, packed=True)
    assert int(it.item()) == int(d["solver_iters"])
    mism = float((c.cpu().numpy() != d["codes"]).mean())
    assert mism < 5e-3, mism
    assert np.abs(s.cpu().numpy().astype(np.float16).astype(np.float32) - d["scale"]).max() <= 1e-6 * np.abs(d["scale"]).max() + 1e-7
    dz = np.abs(z.cpu().numpy() - d["zero"])
    assert float((dz > 1e-3).mean()) < 0.02
    # default of the reference-facing class = the reference's behaviour on a GPU (fp16 solver)
    assert amq.Quantizer.solver_dtype is None
    cfg = amq.BaseQuantizeConfig(nbits=bits, group_size=G)["weight_quant_params"]
    Wq2, meta = amq.Quantizer.quantize(W, device="cuda", compute_dtype=torch.float16, **cfg)
    assert torch.equal(Wq2, W_q)


@pytest.mark.parametrize("bits", [2, 3, 4])
def test_hqq_to_kernel_modules_flow(amq, bits, tmp_path):
    """nn.Linear -> HQQLinear -> prepare_for_inference(gptq / ft) -> forward, plus the cache file:
    second run loads `*_GPTQLinear.pt` / `*_FTLinear.pt` through load_state_dict (patching.py:178-205)."""
    import torch.nn as nn
    torch.manual_seed(bits)
    N, K = 256, 512

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj = nn.Linear(K, N, bias=(bits == 3))

    def make():
        torch.manual_seed(100 + bits)
        blk = Block().half()
        blk.q_proj = amq.HQQLinear(blk.q_proj, amq.BaseQuantizeConfig(nbits=bits, group_size=G), compute_dtype=torch.float16,
                                   device="cuda")
        return blk

    blk = make()
    hq = blk.q_proj
    W_deq = hq.dequantize().float()
    codes_hqq = hq.unpack_codes()
    x = torch.randn(5, K, device="cuda").half()
    y_hqq = hq(x)
    bias = hq.bias.float() if hq.bias is not None else 0.0
    ref = x.float() @ W_deq.t() + bias
    assert O.max_rel(y_hqq.cpu(), ref.cpu()) < 2e-3
    backend = "ft" if bits == 4 else "gptq"
    cache = str(tmp_path / f"m_{bits}bit_128gs_1axis_{'FTLinear' if bits == 4 else 'GPTQLinear'}.pt")
    amq.prepare_for_inference(blk, backend=backend, load_path=cache)
    lin = blk.q_proj
    assert type(lin).__name__ == ("FT_QuantLinear" if bits == 4 else "GPTQLinear") and os.path.exists(cache)
    assert hasattr(lin, "weight")                       # dummy param HF code touches (patching.py:76-91)
    from amq_b200 import ops, _lib
    layout = _lib.LAYOUT_FT if bits == 4 else _lib.LAYOUT_GPTQ
    assert torch.equal(ops.unpack_codes(lin.qweight, bits, layout, N, K, G), codes_hqq)      # codes survive, bit-exact
    for M in (1, 5, 16, 40):                            # decode kernel and prefill entry point
        xm = torch.randn(2, M // 2 if M > 1 else 1, K, device="cuda").half() if M > 1 else torch.randn(1, 1, K, device="cuda").half()
        y = lin(xm)
        assert y.shape == xm.shape[:-1] + (N,) and y.dtype == torch.float16
        r = xm.reshape(-1, K).float() @ W_deq.t() + bias
        assert O.max_rel(y.reshape(-1, N).cpu(), r.cpu()) <= 1e-3
    # second run: empty shells + load_state_dict of the cache file
    blk2 = make()
    amq.prepare_for_inference(blk2, backend=backend, load_path=cache)
    y2 = blk2.q_proj(x)
    assert torch.equal(y2, lin(x))
    # the dummy .weight is added after the cache is written, exactly as in the reference (patching.py:218-222)
    assert set(blk2.state_dict()) - {"q_proj.weight"} == set(torch.load(cache, weights_only=True))


@pytest.mark.parametrize("bits", [2, 3, 4])
def test_hqqlinear_fused_backend(amq, bits):
    """HQQLinear.forward on the search-stage shape (M = 2048 tokens) without materialising the fp16
    weight: W_q is transcoded (codes bit-exact) to the kernel-native layout once and the tcgen05 GEMM / decode GEMV
    run on it.  Must agree with the reference-style backend (dequantise + matmul) and with fp32."""
    import torch.nn as nn
    torch.manual_seed(bits)
    N, K = 512, 1024
    lin = nn.Linear(K, N, bias=True).half()
    hq = amq.HQQLinear(lin, amq.BaseQuantizeConfig(nbits=bits, group_size=G), compute_dtype=torch.float16, device="cuda")
    assert amq.HQQLinear.backend == "fused"
    nat = hq.native_weight()
    from amq_b200 import ops, _lib
    assert nat is not None and nat.numel() == ops.native_bytes(bits, N, K)
    assert torch.equal(ops.unpack_codes(nat, bits, _lib.LAYOUT_NATIVE, N, K, G), hq.unpack_codes())
    assert hq.native_weight() is nat                       # cached until W_q / meta are replaced
    W = hq.dequantize().float()
    for shape in [(1, 2048, K), (3, K), (1, 1, K)]:
        x = torch.randn(*shape, device="cuda").half()
        y = hq(x)
        ref = x.float() @ W.t() + hq.bias.float()
        assert y.shape == x.shape[:-1] + (N,) and y.dtype == torch.float16
        assert O.max_rel(y.reshape(-1, N).cpu(), ref.reshape(-1, N).cpu()) <= 1e-3
        # the reference's own form: x @ dequantize().T + bias in fp16 (quantize.py:893-898), computed here by the test
        y_pt = torch.matmul(x, hq.dequantize().t()) + hq.bias
        assert O.max_rel(y.reshape(-1, N).cpu(), y_pt.reshape(-1, N).float().cpu()) <= 2e-3
    amq.HQQLinear.set_backend("pytorch")                   # reference backend names are accepted and map onto the fused kernels
    assert amq.HQQLinear.backend == "fused"
    with pytest.raises(ValueError):
        amq.HQQLinear.set_backend("no-such-backend")
    # a shape the native record grid does not take (N % 32 != 0): converted once to the GPTQ layout, any-shape kernel
    hq48 = amq.HQQLinear(nn.Linear(256, 48, bias=False).half(), amq.BaseQuantizeConfig(nbits=bits, group_size=G),
                         compute_dtype=torch.float16, device="cuda")
    assert hq48.native_weight() is None
    x = torch.randn(4, 256, device="cuda").half()
    assert O.max_rel(hq48(x).cpu(), (x.float() @ hq48.dequantize().float().t()).cpu()) <= 2e-3


def test_gptq_module_on_golden_reference_buffers(amq):
    """A GPTQLinear filled with the buffers the REFERENCE produced (golden) reproduces the oracle."""
    d = np.load(os.path.join(GOLD, "linear_3bit_N256_K512.npz"))
    N, K = d["W"].shape
    m = amq.GPTQLinear(3, G, K, N, bias=False)
    m.load_state_dict({"qweight": torch.from_numpy(d["gptq_qweight"]), "scales": torch.from_numpy(d["gptq_scales"]),
                       "zeros": torch.from_numpy(d["gptq_zeros"])})
    m = m.cuda()
    for M in (1, 5):
        y = m(torch.from_numpy(d[f"x{M}"]).cuda())
        assert O.max_rel(y.cpu(), torch.from_numpy(d[f"y{M}_fp32"])) <= 1e-3
        assert O.max_rel(y.cpu(), torch.from_numpy(d[f"y{M}_ref_fp16"])) <= 2e-3
    # pack() on the GPU reproduces the reference's packed buffers from the dequantised weights
    m2 = amq.GPTQLinear(3, G, K, N, bias=False).cuda()
    s = torch.from_numpy(d["hqq_scale"]).half().reshape(N, -1).cuda()
    z = torch.from_numpy(d["hqq_zero"]).half().reshape(N, -1).cuda()
    m2.pack(torch.from_numpy(d["W_deq"]).cuda(), s, z)
    assert np.array_equal(m2.qweight.cpu().numpy(), d["gptq_qweight"])


def test_pack_intweight_matches_reference_layout(amq):
    d = np.load(os.path.join(GOLD, "linear_4bit_N256_K512.npz"))
    codes = torch.from_numpy(d["codes"].reshape(256, 512).astype(np.int32)).cuda()
    q = amq.pack_intweight(codes, interleave=4, kstride=64)
    assert np.array_equal(q.cpu().numpy(), d["ft_qweight"])


@pytest.mark.parametrize("bits", [2, 3, 4])
def test_config4_proxy_sweep_shapes(amq, bits):
    """Config 4 at Qwen2-7B shapes (k/v 512x3584 with the full K, one q-sized slab): quantize -> pack ->
    dequantize round trip properties at full width: codes in range, pack/unpack bit-exact, error bound."""
    torch.manual_seed(0)
    N, K = 512, 3584
    W = (torch.randn(N, K, device="cuda") * 0.02).half()
    cfg = amq.BaseQuantizeConfig(nbits=bits, group_size=G)["weight_quant_params"]
    amq.Quantizer.solver_dtype = torch.float32        # the oracle-pinned arithmetic (the default follows the reference on a GPU: fp16)
    try:
        W_q, meta = amq.Quantizer.quantize(W, device="cuda", compute_dtype=torch.float16, **cfg)
    finally:
        amq.Quantizer.solver_dtype = None
    R = N * K // G
    codes = amq.Quantizer.unpack[meta["packing"]](W_q)[:R]
    assert int(codes.max()) <= 2 ** bits - 1
    assert torch.equal(amq.Quantizer.pack[meta["packing"]](codes), W_q)
    nt = torch.get_num_threads()
    torch.set_num_threads(1)                         # torch's vectorised pow for every element (no scalar chunk tails)
    try:
        o_codes, o_scale, o_zero, _ = O.hqq_quantize(W.cpu(), bits, G)
    finally:
        torch.set_num_threads(nt)
    assert np.array_equal(codes.cpu().numpy(), o_codes)
    assert torch.equal(meta["scale"].cpu(), o_scale) and torch.equal(meta["zero"].cpu(), o_zero)
    meta16 = dict(meta, scale=meta["scale"].half(), zero=meta["zero"].half())
    W_r = amq.Quantizer.dequantize(W_q, meta16)
    step = meta["scale"].max().item()
    assert float((W_r.float() - W.float()).abs().max()) <= 1.01 * step * (1.0 if bits > 2 else 1.5)
