"""Round-2 parity gaps: adversarial activations for the decode GEMV / prefill GEMM (the IMMA path turns every
128-group of x into fixed point relative to the group's largest exponent, csrc/gemv_mma.cuh `item_stats`), the
non-finite policy, fp32-meta call paths of the Python shims, reference-written HQQLinear state dicts, the 8-bit
GPTQLinear constructor path and per-device launches.  Tolerance: max|y - ref| / max|ref| <= 1e-3 against the exact
fp32 reference from the stored buffers."""
import os

import numpy as np
import pytest
import torch

from oracle import amq_oracle as O

pytestmark = pytest.mark.gpu

G = 128
TOL = 1e-3
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import amq_b200.ops as ops_
    return ops_


@pytest.fixture(scope="module")
def amq():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import amq_b200
    return amq_b200


def _layer(ops, N, K, bits, seed):
    rs = np.random.RandomState(seed)
    codes = torch.from_numpy(rs.randint(0, 2 ** bits, size=(N, K)).astype(np.uint8)).cuda()
    lo, hi = {2: (0.024, 0.054), 3: (0.010, 0.023), 4: (0.0047, 0.011)}[bits]
    scale = torch.from_numpy(rs.uniform(lo, hi, size=(N, K // G)).astype(np.float32)).half().cuda()
    zero = torch.from_numpy(rs.uniform(0.5, 2 ** bits - 1.5, size=(N, K // G)).astype(np.float32)).half().cuda()
    nat = ops.pack_native(bits, codes, scale, zero)
    zs = (zero * scale).float()                                     # fp16 product, as autogptq.py:112
    W = (codes.float().reshape(N, K // G, G) * scale.float()[..., None] - zs[..., None]).reshape(N, K)
    return nat, W


def _adversarial(kind, M, K, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g)
    n_g = K // G
    if kind == "outlier_per_group":          # one element 2^10 times larger than the rest in EVERY group
        idx = torch.randint(0, G, (M, n_g), generator=g)
        xv = x.reshape(M, n_g, G)
        xv.scatter_(2, idx[..., None], xv.gather(2, idx[..., None]) * 1024.0)
    elif kind == "outlier_some_groups":      # a few huge channels (the residual-stream pattern of Llama), rest O(1)
        cols = torch.randperm(K, generator=g)[:6]
        x[:, cols] *= 2000.0
    elif kind == "zero_groups":              # whole groups exactly zero (group max exponent 0), others normal
        xv = x.reshape(M, n_g, G)
        xv[:, ::2] = 0.0
    elif kind == "all_zero":
        x.zero_()
    elif kind == "subnormal":                # every element below fp16's smallest normal (2^-14)
        x = x * 2.0 ** -18
    elif kind == "mixed_subnormal":          # one normal element per group, the rest subnormal
        x = x * 2.0 ** -18
        x.reshape(M, n_g, G)[:, :, 5] = 0.37
    elif kind == "max_magnitude":            # +-65504 entries: sparse, so that y stays inside fp16
        x = x * 0.01
        cols = torch.randperm(K, generator=g)[:4]
        x[:, cols] = torch.tensor([65504.0, -65504.0, 65504.0, -65504.0])
    elif kind == "tiny_and_huge":            # exponents 1 and 30 in the same group
        x = x * 2.0 ** -13
        x.reshape(M, n_g, G)[:, :, 7] = 30000.0
    else:
        raise ValueError(kind)
    return x.half()


KINDS = ["outlier_per_group", "outlier_some_groups", "zero_groups", "all_zero", "subnormal", "mixed_subnormal",
         "max_magnitude", "tiny_and_huge"]


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("kind", KINDS)
def test_gemv_adversarial_activations(ops, bits, kind):
    N, K = 512, 1024
    nat, W = _layer(ops, N, K, bits, seed=bits)
    for M in (1, 2, 5, 16):
        x = _adversarial(kind, M, K, seed=10 * bits + M).cuda()
        y = ops.gemv(bits, nat, x, N, K)
        ref = x.float() @ W.t()
        assert torch.isfinite(y).all(), (bits, kind, M)
        if kind == "all_zero":
            assert (y == 0).all()
            continue
        # outputs in fp16's subnormal range are quantised to multiples of 2^-24 whatever the kernel does: the bound is
        # TOL relative to the largest output, plus half of that step
        err = float((y.float() - ref).abs().max())
        assert err <= TOL * float(ref.abs().max()) + 2.0 ** -25, (bits, kind, M, err, float(ref.abs().max()))
        if kind == "outlier_some_groups":
            # the groups WITHOUT an outlier must keep their own precision (per-group exponent): check the rows' error
            # against the magnitude the quiet groups alone produce, not against the outlier-dominated maximum
            quiet = x.clone().float()
            quiet[quiet.abs() > 100] = 0
            yq = ops.gemv(bits, nat, quiet.half(), N, K)
            assert O.max_rel(yq.float().cpu(), (quiet @ W.t()).cpu()) <= TOL


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("kind", ["outlier_per_group", "outlier_some_groups", "zero_groups", "mixed_subnormal", "max_magnitude"])
def test_prefill_gemm_adversarial_activations(ops, bits, kind):
    N, K, M = 512, 1024, 96
    nat, W = _layer(ops, N, K, bits, seed=20 + bits)
    x = _adversarial(kind, M, K, seed=77 + bits).cuda()
    y = ops.gemm_tc(bits, nat, x, N, K)
    ref = x.float() @ W.t()
    assert torch.isfinite(y).all()
    err = float((y.float() - ref).abs().max())
    assert err <= TOL * float(ref.abs().max()) + 2.0 ** -25, (bits, kind, err, float(ref.abs().max()))


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("bad", [float("inf"), float("-inf"), float("nan")])
def test_nonfinite_activation_policy(ops, bits, bad):
    """Policy: an activation row containing inf / NaN yields NaN for EVERY output of that row (the reference's
    fp16 matmul yields +-inf or NaN there; an integer dot product cannot carry inf, so the row is poisoned instead of
    returning finite garbage); other rows of the batch are unaffected."""
    N, K = 256, 512
    nat, W = _layer(ops, N, K, bits, seed=30 + bits)
    for M in (1, 2, 6):
        x = torch.randn(M, K).half()
        x[M - 1, 300] = bad
        y = ops.gemv(bits, nat, x.cuda(), N, K)
        assert torch.isnan(y[M - 1]).all(), (bits, bad, M)
        if M > 1:
            ref = x[: M - 1].float().cuda() @ W.t()
            assert O.max_rel(y[: M - 1].float().cpu(), ref.cpu()) <= TOL
    xm = torch.randn(80, K).half()
    xm[3, 17] = bad
    ym = ops.gemm_tc(bits, nat, xm.cuda(), N, K)
    assert not torch.isfinite(ym[3]).any()                 # tcgen05 fp16 MMA propagates inf / NaN natively
    ok = torch.ones(80, dtype=torch.bool)
    ok[3] = False
    assert torch.isfinite(ym[ok.cuda()]).all()


def test_fp32_meta_goes_through_the_shims(ops, amq):
    """dtype-conversion temporaries must outlive the launch.  Quantizer.quantize returns fp32 scale / zero;
    dequantising with them directly (two same-sized temporaries -> the caching allocator would alias them) must equal
    dequantising with pre-cast fp16 meta.  Same for GPTQLinear.pack with fp32 scales / zeros and repack after .half()."""
    torch.manual_seed(0)
    N, K = 256, 512
    W = (torch.randn(N, K) * 0.02).half().cuda()
    for bits in (2, 3, 4):
        W_q, meta = amq.Quantizer.quantize(W, nbits=bits, group_size=G, axis=1, round_zero=(bits == 4))
        assert meta["scale"].dtype == torch.float32
        d32 = amq.Quantizer.dequantize(W_q, dict(meta))
        d16 = amq.Quantizer.dequantize(W_q, dict(meta, scale=meta["scale"].half(), zero=meta["zero"].half()))
        assert torch.equal(d32, d16)
        s, z = meta["scale"].reshape(N, -1), meta["zero"].reshape(N, -1)
        q32 = ops.gptq_pack(bits, d16, s, z, G)                       # fp32 scales / zeros
        q16 = ops.gptq_pack(bits, d16, s.half(), z.half(), G)
        assert all(torch.equal(a, b) for a, b in zip(q32, q16))
        m = amq.GPTQLinear(bits, G, K, N, bias=False).cuda()
        m.pack(d16, s, z)
        x = torch.randn(3, K).half().cuda()
        y = m(x)
        m2 = amq.GPTQLinear(bits, G, K, N, bias=False).cuda()
        m2.load_state_dict(m.state_dict())
        m2 = m2.half()                                                # fp32 buffers -> fp16: repack converts them back
        assert torch.equal(m2(x), y)


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("form", ["encoded", "plain"])
def test_hqqlinear_loads_reference_written_state_dict(amq, bits, form):
    """tests/golden/hqqlinear_state_*.pt hold state dicts written by the REFERENCE's HQQLinear.state_dict() (encoded =
    its default, plain = what qmodel.pt stores) and the reference's forward on them (oracle/gen_golden.py)."""
    d = torch.load(os.path.join(GOLD, f"hqqlinear_state_{bits}bit.pt"), weights_only=True)
    layer = amq.HQQLinear(None, None, compute_dtype=torch.float16, device="cuda")
    layer.load_state_dict(d[form])
    assert layer.meta["nbits"] == bits and layer.meta["group_size"] == G and tuple(layer.meta["shape"]) == (64, 256)
    assert layer.meta["packing"] == {2: "2bit_u8", 3: "3bit_32", 4: "4bit_u8"}[bits]
    assert torch.equal(layer.dequantize().cpu(), d["W_deq"])          # two-rounding fp16 dequant, bit-exact
    x = d["x"].cuda()
    y = layer(x)
    assert O.max_rel(y.float().cpu(), d["y_fp32"]) <= TOL
    assert O.max_rel(y.float().cpu(), d["y_ref_fp16"].float()) <= 2e-3
    # round trip: our state dict carries the same keys and tensors
    sd = layer.state_dict()
    for k, v in d["plain"].items():
        assert k in sd, k
        if isinstance(v, torch.Tensor):
            assert torch.equal(sd[k].detach().cpu(), v), k

    # through a PARENT module's load_state_dict (quantize.py:684-706 `_load_from_state_dict`)
    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.proj = amq.HQQLinear(None, None, compute_dtype=torch.float16, device="cuda")
            self.norm = torch.nn.LayerNorm(8)

    blk = Block()
    flat = {"proj." + k: v for k, v in d[form].items()}
    flat.update({"norm." + k: v for k, v in blk.norm.state_dict().items()})
    blk.load_state_dict(flat, strict=True)
    assert torch.equal(blk.proj(x), y)


def test_gptqlinear_8bit_constructor_path(amq):
    """autogptq.py:43-46,86: the reference constructor accepts bits = 8 (use_cuda_fp16 forced off); its small-M branch
    then raises, only the large-M branch (:245-283) serves such a module.  AMQ never builds one."""
    N, K = 64, 256
    m = amq.GPTQLinear(8, G, K, N, bias=False).cuda()
    assert m.qweight.shape == (K // 32 * 8, N) and m.use_cuda_fp16 is False and m.maxq == 255
    rs = np.random.RandomState(0)
    codes = rs.randint(0, 256, size=(N, K))
    scale = torch.from_numpy(rs.uniform(0.0005, 0.001, size=(N, K // G)).astype(np.float32)).half()
    zero = torch.from_numpy(rs.uniform(100, 150, size=(N, K // G)).astype(np.float32)).half()
    W = ((torch.from_numpy(codes).float().reshape(N, K // G, G) - zero.float()[..., None]) * scale.float()[..., None]).reshape(N, K).half()
    m.pack(W.cuda(), scale.cuda(), zero.cuda())
    qw = m.qweight.cpu().numpy()
    got = O.gptq_unpack_fast(qw, 8)                                   # [K, N]
    ref_codes = O.gptq_codes_from_weight(W, scale, zero, G)
    assert np.array_equal(got, ref_codes.T.astype(np.uint8))
    x = torch.randn(130, K).half().cuda()
    y = m(x)
    ref = O.gptq_forward_fp32(x.cpu(), qw, m.scales.cpu(), m.zeros.cpu(), 8, G)
    assert O.max_rel(y.float().cpu(), ref) <= TOL
    with pytest.raises(NotImplementedError):
        m(x[:4])


def test_ops_follow_the_tensors_device(ops, amq):
    """launches must go to the device (and stream, workspace, per-device kernel attributes) of the tensors,
    not of whatever device is current.  Needs two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    N, K, bits = 256, 512, 3
    torch.cuda.set_device(0)
    with torch.cuda.device(1):
        nat, W = _layer(ops, N, K, bits, seed=5)
        nat, W = nat.to("cuda:1"), W.to("cuda:1")
    x = torch.randn(2, K).half().to("cuda:1")
    assert torch.cuda.current_device() == 0
    y = ops.gemv(bits, nat, x, N, K)
    assert y.device == x.device and O.max_rel(y.float().cpu(), (x.float() @ W.t()).cpu()) <= TOL
    xl = torch.randn(64, K).half().to("cuda:1")
    yl = ops.gemm_tc(bits, nat, xl, N, K)
    assert O.max_rel(yl.float().cpu(), (xl.float() @ W.t()).cpu()) <= TOL
