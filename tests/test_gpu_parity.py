"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed
golden fixtures made from the unmodified reference.  Integer / layout work must be bit-exact;
outputs must satisfy max|y - ref| / max|ref| <= 1e-3 against the exact-fp32 reference computed
from the stored buffers."""
import glob
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


def _synthetic(N, K, bits, seed):
    """Fast synthetic linear: random codes, realistic fp16 scale / fractional zero."""
    rs = np.random.RandomState(seed)
    codes = rs.randint(0, 2 ** bits, size=(N, K)).astype(np.uint8)
    lo, hi = {2: (0.024, 0.054), 3: (0.010, 0.023), 4: (0.0047, 0.011)}[bits]
    scale = torch.from_numpy(rs.uniform(lo, hi, size=(N, K // G)).astype(np.float32)).half()
    zero = torch.from_numpy(rs.uniform(0.5, 2 ** bits - 1.5, size=(N, K // G)).astype(np.float32)).half()
    return codes, scale, zero


def _gptq_buffers(codes, scale, zero, bits):
    qweight = O.gptq_pack_codes(codes.astype(np.int64), bits)
    scales = scale.t().contiguous().float()
    zeros = (zero * scale).t().contiguous().float()        # fp16 product, as autogptq.py:112
    return qweight, scales, zeros


# ------------------------------------------------------------------ codes: bit-exact
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "linear_*.npz"))))
def test_unpack_all_layouts_golden(ops, path):
    from amq_b200 import _lib
    d = np.load(path)
    bits = int(os.path.basename(path).split("_")[1][0])
    N, K = d["W"].shape
    codes = d["codes"].reshape(N, K)
    dev = "cuda"
    got = ops.unpack_codes(torch.from_numpy(d["hqq_Wq"]).to(dev), bits, _lib.LAYOUT_HQQ, N, K, G).cpu().numpy()
    assert np.array_equal(got, codes)
    got = ops.unpack_codes(torch.from_numpy(d["gptq_qweight"]).to(dev), bits, _lib.LAYOUT_GPTQ, N, K, G).cpu().numpy()
    assert np.array_equal(got, codes)
    if bits == 4:
        got = ops.unpack_codes(torch.from_numpy(d["ft_qweight"]).to(dev), bits, _lib.LAYOUT_FT, N, K, G).cpu().numpy()
        assert np.array_equal(got, codes)
    nat = ops.repack_gptq(bits, torch.from_numpy(d["gptq_qweight"]).to(dev), torch.from_numpy(d["gptq_scales"]).to(dev),
                          torch.from_numpy(d["gptq_zeros"]).to(dev), N, K, G)
    got = ops.unpack_codes(nat, bits, _lib.LAYOUT_NATIVE, N, K, G).cpu().numpy()
    assert np.array_equal(got, codes)


@pytest.mark.parametrize("bits", [2, 3, 4])
def test_packers_match_reference_golden(ops, bits):
    """GPTQLinear.pack / FT pack / BitPack on the GPU reproduce the reference's packed tensors."""
    d = np.load(os.path.join(GOLD, f"linear_{bits}bit_N256_K512.npz"))
    N, K = d["W"].shape
    dev = "cuda"
    W_deq = torch.from_numpy(d["W_deq"]).to(dev)
    s = torch.from_numpy(d["hqq_scale"]).half().reshape(N, -1).to(dev)
    z = torch.from_numpy(d["hqq_zero"]).half().reshape(N, -1).to(dev)
    q, so, zo = ops.gptq_pack(bits, W_deq, s, z, G)
    assert np.array_equal(q.cpu().numpy(), d["gptq_qweight"])
    assert np.array_equal(so.cpu().numpy(), d["gptq_scales"]) and np.array_equal(zo.cpu().numpy(), d["gptq_zeros"])
    codes = torch.from_numpy(d["codes"]).to(dev)
    assert np.array_equal(ops.hqq_pack(bits, codes).cpu().numpy(), d["hqq_Wq"])
    assert np.array_equal(ops.hqq_unpack(bits, torch.from_numpy(d["hqq_Wq"]).to(dev), codes.shape[0]).cpu().numpy(), d["codes"])
    deq = ops.hqq_dequant(bits, torch.from_numpy(d["hqq_Wq"]).to(dev), s.reshape(-1), z.reshape(-1), N, K, G)
    assert np.array_equal(deq.cpu().numpy(), d["W_deq"])       # two fp16 roundings, bit-exact
    if bits == 4:
        fq, fs, fz = ops.ft_pack(W_deq, s, z, G)
        assert np.array_equal(fq.cpu().numpy(), d["ft_qweight"])
        assert np.array_equal(fs.cpu().numpy(), d["gptq_scales"].astype(np.float16))
        assert np.array_equal(fz.cpu().numpy(), (-d["gptq_zeros"]).astype(np.float16))


@pytest.mark.parametrize("R", [40, 130, 1310])
def test_bitpack_roundtrip_property(ops, R):
    """The reference's tests/test_bitpack.py property: unpack(pack(W)) == W, bit-exact, incl. 3-bit padding."""
    torch.manual_seed(42)
    for bits in (2, 3, 4):
        if bits != 3 and R % (2 if bits == 4 else 4):
            continue
        codes = torch.randint(0, 2 ** bits, (R, G), dtype=torch.uint8, device="cuda")
        packed = ops.hqq_pack(bits, codes)
        assert np.array_equal(packed.cpu().numpy(), O.hqq_pack(codes.cpu().numpy(), bits))
        assert torch.equal(ops.hqq_unpack(bits, packed, R), codes)


# ------------------------------------------------------------------ outputs: <= 1e-3 max-rel vs fp32
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "linear_*.npz"))))
def test_gemv_golden(ops, path):
    d = np.load(path)
    bits = int(os.path.basename(path).split("_")[1][0])
    N, K = d["W"].shape
    dev = "cuda"
    qw = torch.from_numpy(d["gptq_qweight"]).to(dev)
    sc = torch.from_numpy(d["gptq_scales"]).to(dev)
    ze = torch.from_numpy(d["gptq_zeros"]).to(dev)
    nat = ops.repack_gptq(bits, qw, sc, ze, N, K, G)
    for M in (1, 5):
        x = torch.from_numpy(d[f"x{M}"]).to(dev)
        ref = torch.from_numpy(d[f"y{M}_fp32"])
        y = ops.gemv(bits, nat, x, N, K).cpu()
        assert O.max_rel(y, ref) <= TOL, (bits, M, O.max_rel(y, ref))
        y2 = ops.gemv_gptq_layout(bits, qw, sc, ze, x, N, K, G).cpu()
        assert O.max_rel(y2, ref) <= TOL
        # distance to the reference's own fp16 torch output, reported for context
        assert O.max_rel(y, torch.from_numpy(d[f"y{M}_ref_fp16"])) <= 2e-3


SHAPES = [(4096, 4096), (11008, 4096), (4096, 11008), (1024, 4096), (128, 8192), (512, 3584), (3584, 18944), (32, 128)]


@pytest.mark.parametrize("N,K", SHAPES)
@pytest.mark.parametrize("bits", [2, 3, 4])
def test_gemv_model_shapes(ops, N, K, bits):
    """Full model shapes (configs 1-3, 5 and the tp=8 shards): GPU fp32 torch reference from the same
    codes (the CPU oracle is checked against it on the 4096x4096 case below)."""
    from amq_b200 import _lib
    dev = "cuda"
    codes, scale, zero = _synthetic(N, K, bits, seed=N + K + bits)
    cg = torch.from_numpy(codes).to(dev)
    sg, zg = scale.to(dev), zero.to(dev)
    nat = ops.pack_native(bits, cg, sg, zg)
    assert torch.equal(ops.unpack_codes(nat, bits, _lib.LAYOUT_NATIVE, N, K, G), cg)
    zs = (zg * sg)                                            # fp16 product
    W = (cg.float().reshape(N, K // G, G) * sg.float()[..., None] - zs.float()[..., None]).reshape(N, K)
    torch.manual_seed(bits)
    bias = torch.randn(N, device=dev).half()
    for M in (1, 2, 8, 16):
        x = torch.randn(M, K, device=dev).half()
        ref = x.float() @ W.t()
        y = ops.gemv(bits, nat, x, N, K)
        assert O.max_rel(y.cpu(), ref.cpu()) <= TOL, (N, K, bits, M, O.max_rel(y.cpu(), ref.cpu()))
        if M in (1, 16):
            yb = ops.gemv(bits, nat, x, N, K, bias)
            assert O.max_rel(yb.cpu(), (ref + bias.float()).cpu()) <= TOL
    # determinism: no atomics on data -> bitwise identical reruns
    x = torch.randn(1, K, device=dev).half()
    y1 = ops.gemv(bits, nat, x, N, K)
    y2 = ops.gemv(bits, nat, x, N, K)
    assert torch.equal(y1, y2)


@pytest.mark.parametrize("widths", [(2,), (3,), (4,), (2, 3), (2, 4), (3, 4), (2, 3, 4), (4, 2, 4), (3, 3, 2)])
@pytest.mark.parametrize("pro", ["none", "rmsnorm"])
def test_grouped_launch_every_bit_width_set(ops, widths, pro):
    """Batch-1 grouped launches are served by a kernel instance per SET of bit widths present (one width, two widths, all
    three): every set, with and without the RMSNorm prologue, against the fp32 statement of the same problems."""
    from amq_b200 import _lib
    dev = torch.device("cuda")
    H, N = 1024, 2560            # 80 row blocks: too many for a cluster K split, so these are the slim per-width-set instances
    torch.manual_seed(sum(widths))
    x = torch.randn(1, H, device=dev).half()
    gamma = (1.0 + 0.1 * torch.randn(H, device=dev)).half()
    eps = 1e-5
    xin = x.float()
    if pro == "rmsnorm":
        xin = (gamma.float() * (xin * torch.rsqrt(xin.pow(2).mean(-1, keepdim=True) + eps)).half().float())
    y = torch.zeros(1, N * len(widths), device=dev, dtype=torch.float16)
    ws = ops.workspace(dev, N * len(widths), H, 1)
    probs, refs = [], []
    for i, b in enumerate(widths):
        codes, scale, zero = _synthetic(N, H, b, seed=100 + 7 * i + b)
        cg, sg, zg = torch.from_numpy(codes).to(dev), scale.to(dev), zero.to(dev)
        nat = ops.pack_native(b, cg, sg, zg)
        W = (cg.float().reshape(N, H // G, G) * sg.float()[..., None] - (zg * sg).float()[..., None]).reshape(N, H)
        refs.append(xin @ W.t())
        p = ops.make_problem(b, nat, x, y, N, H, ldy=N * len(widths),
                             prologue=_lib.PRO_RMSNORM if pro == "rmsnorm" else _lib.PRO_NONE,
                             gamma=gamma if pro == "rmsnorm" else None, eps=eps)
        p.y = y.data_ptr() + 2 * N * i
        probs.append((p, nat))
    ops.gemv_grouped([p for p, _ in probs], ws)
    torch.cuda.synchronize()
    ref = torch.cat(refs, dim=1)
    tol = TOL if pro == "none" else 2e-3                      # the prologue rounds x * rs to fp16 before gamma (HF's RMSNorm)
    assert O.max_rel(y.cpu(), ref.cpu()) <= tol, (widths, pro, O.max_rel(y.cpu(), ref.cpu()))


@pytest.mark.parametrize("M", [1])
def test_output_activation_and_mul_prologue(ops, M):
    """gate|up -> down as QuantDecoder launches it: `act = 1` on the gate problem stores silu(gate) and the down launch reads
    x = a * b (AMQB_PRO_MUL) - bit for bit what the one-sided AMQB_PRO_SILU_MUL prologue computes from raw gate / up, and
    within tolerance of the fp32 statement silu(g) * u."""
    from amq_b200 import _lib
    dev = "cuda"
    H, I, bits = 512, 1024, 3
    cg, sg, zg = [t if isinstance(t, torch.Tensor) else torch.from_numpy(t) for t in _synthetic(I, H, bits, seed=11)]
    cu, su, zu = [t if isinstance(t, torch.Tensor) else torch.from_numpy(t) for t in _synthetic(I, H, bits, seed=12)]
    cd, sd, zd = [t if isinstance(t, torch.Tensor) else torch.from_numpy(t) for t in _synthetic(H, I, bits, seed=13)]
    wg = ops.pack_native(bits, cg.to(dev), sg.to(dev), zg.to(dev))
    wu = ops.pack_native(bits, cu.to(dev), su.to(dev), zu.to(dev))
    wd = ops.pack_native(bits, cd.to(dev), sd.to(dev), zd.to(dev))
    torch.manual_seed(M)
    x = torch.randn(M, H, device=dev).half()
    ws = ops.workspace(x.device, 2 * I, max(H, I), M)
    outs = []
    for act_in_producer in (False, True):
        gu = torch.zeros(M, 2 * I, device=dev, dtype=torch.float16)
        y = torch.zeros(M, H, device=dev, dtype=torch.float16)
        pg = ops.make_problem(bits, wg, x, gu, I, H, ldy=2 * I)
        pu = ops.make_problem(bits, wu, x, gu, I, H, ldy=2 * I)
        pu.y = gu.data_ptr() + 2 * I
        pg.act = int(act_in_producer)
        ops.gemv_grouped([pg, pu], ws)
        pd = ops.make_problem(bits, wd, gu, y, H, I, ldx=2 * I,
                              prologue=_lib.PRO_MUL if act_in_producer else _lib.PRO_SILU_MUL)
        ops.gemv_grouped([pd], ws)
        torch.cuda.synchronize()
        outs.append((gu.clone(), y.clone()))
    (gu0, y0), (gu1, y1) = outs
    assert torch.equal(gu0[:, I:], gu1[:, I:])                                    # up half untouched by the activation
    g = gu0[:, :I].float()
    assert torch.equal(gu1[:, :I], (g / (1 + torch.exp(-g))).half()) or \
        O.max_rel(gu1[:, :I].cpu(), (g / (1 + torch.exp(-g))).cpu()) <= 1e-3      # __expf / __fdividef vs torch
    assert torch.equal(y0, y1)                                                    # same arithmetic on either side
    Wd = (cd.float().reshape(H, I // G, G) * sd.float()[..., None] - (zd * sd).float()[..., None]).reshape(H, I).to(dev)
    h = (torch.nn.functional.silu(gu0[:, :I].float()).half() * gu0[:, I:]).float()
    assert O.max_rel(y1.cpu(), (h @ Wd.t()).cpu()) <= TOL


@pytest.mark.parametrize("bits", [2, 3, 4])
def test_gemv_group_count_sweep(ops, bits, monkeypatch):
    """Every k-group count around the pipeline-stage boundaries (a stage is 16 records; the batch-1 / batch-2 consumers
    take two stages per iteration: 1, 15..17, 31..33, 47..49 groups exercise the single-stage tail, the half-filled
    second stage and the odd stage count), both row-block parities, with and without the two-stage pairing
    (AMQB_NO_PAIR): the pairing must not change a single bit (same per-warp accumulation order)."""
    dev = "cuda"
    for n_g in (1, 2, 15, 16, 17, 31, 32, 33, 47, 48, 49, 64, 86):
        K = 128 * n_g
        for N in (32, 96):
            codes, scale, zero = _synthetic(N, K, bits, seed=n_g + N)
            cg = torch.from_numpy(codes).to(dev)
            sg, zg = scale.to(dev), zero.to(dev)
            nat = ops.pack_native(bits, cg, sg, zg)
            W = (cg.float().reshape(N, K // G, G) * sg.float()[..., None] - (zg * sg).float()[..., None]).reshape(N, K)
            for M in (1, 2):
                x = torch.randn(M, K, device=dev).half()
                ref = x.float() @ W.t()
                monkeypatch.delenv("AMQB_NO_PAIR", raising=False)
                y = ops.gemv(bits, nat, x, N, K)
                assert O.max_rel(y.cpu(), ref.cpu()) <= TOL, (bits, n_g, N, M, O.max_rel(y.cpu(), ref.cpu()))
                monkeypatch.setenv("AMQB_NO_PAIR", "1")
                y1 = ops.gemv(bits, nat, x, N, K)
                torch.cuda.synchronize()
                assert torch.equal(y, y1), (bits, n_g, N, M)
    monkeypatch.delenv("AMQB_NO_PAIR", raising=False)


@pytest.mark.parametrize("bits", [2, 3, 4])
def test_config1_vs_cpu_oracle(ops, bits):
    """Config 1: 4096x4096 q_proj-shaped linear, batch-1, against the CPU oracle end to end
    (oracle pack -> our repack -> our GEMV vs oracle fp32 and the oracle's fp16 torch path)."""
    N = K = 4096
    codes, scale, zero = _synthetic(N, K, bits, seed=bits)
    qweight, scales, zeros = _gptq_buffers(codes, scale, zero, bits)
    dev = "cuda"
    nat = ops.repack_gptq(bits, torch.from_numpy(qweight).to(dev), scales.to(dev), zeros.to(dev), N, K, G)
    torch.manual_seed(0)
    x = torch.randn(1, K).half()
    ref32 = O.gptq_forward_fp32(x, qweight, scales, zeros, bits, G)
    y = ops.gemv(bits, nat, x.to(dev), N, K).cpu()
    assert O.max_rel(y, ref32) <= TOL
    ref16 = O.gptq_forward_torch(x, qweight, scales, zeros, bits, G)
    assert O.max_rel(y, ref16) <= 2e-3


# ------------------------------------------------------------------ prefill: tcgen05 / TMEM kernel
PREFILL = [(128, 128, 128), (128, 256, 64), (256, 512, 40), (512, 1024, 300), (4096, 4096, 512), (11008, 4096, 130),
           (4096, 11008, 128), (96, 256, 33)]


@pytest.mark.parametrize("N,K,M", PREFILL)
@pytest.mark.parametrize("bits", [2, 3, 4])
def test_prefill_gemm(ops, N, K, M, bits):
    """amqb_gemm_tc (replaces FT gemm_4bit and the large-M torch branch of GPTQLinear.forward) against the exact fp32
    reference from the same codes; ragged M (rows past M are zero-filled tiles), N % 128 != 0 takes the slab path."""
    dev = "cuda"
    codes, scale, zero = _synthetic(N, K, bits, seed=N + K + M + bits)
    cg = torch.from_numpy(codes).to(dev)
    sg, zg = scale.to(dev), zero.to(dev)
    nat = ops.pack_native(bits, cg, sg, zg)
    W = (cg.float().reshape(N, K // G, G) * sg.float()[..., None] - (zg * sg).float()[..., None]).reshape(N, K)
    torch.manual_seed(M)
    x = torch.randn(M, K, device=dev).half()
    bias = torch.randn(N, device=dev).half()
    ref = x.float() @ W.t()
    y = ops.gemm_tc(bits, nat, x, N, K)
    assert O.max_rel(y.cpu(), ref.cpu()) <= TOL, (N, K, M, bits, O.max_rel(y.cpu(), ref.cpu()))
    ws = ops.gemm_workspace(M, K, bits, x.device)
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    yb = ops.gemm_tc(bits, nat, x, N, K, bias, out=out, workspace=ws)
    assert yb.data_ptr() == out.data_ptr()
    assert O.max_rel(yb.cpu(), (ref + bias.float()).cpu()) <= TOL
    assert torch.equal(ops.gemm_tc(bits, nat, x, N, K, bias), yb)          # deterministic
    # the same rows through the decode kernel (M <= 16): both paths see the same weights
    y16 = ops.gemv(bits, nat, x[:16].contiguous(), N, K)
    assert O.max_rel(y16.cpu(), ref[:16].cpu()) <= TOL


@pytest.mark.parametrize("M", [40, 300])
def test_prefill_gemm_grouped_matches_single_calls(ops, M):
    """q|k|v-style grouped tcgen05 launch (three problems of different width and size over the same activations, with and
    without bias) against the same problems through amqb_gemm_tc one by one: bit-identical (same tiles, same arithmetic)."""
    dev = torch.device("cuda")
    K = 1024
    torch.manual_seed(M)
    x = torch.randn(M, K, device=dev).half()
    members, singles = [], []
    for i, (bits, N) in enumerate(((3, 512), (2, 128), (4, 256))):
        codes, scale, zero = _synthetic(N, K, bits, seed=40 + i)
        nat = ops.pack_native(bits, torch.from_numpy(codes).to(dev), scale.to(dev), zero.to(dev))
        bias = torch.randn(N, device=dev).half() if i != 1 else None
        members.append((bits, nat, N, bias))
        singles.append(ops.gemm_tc(bits, nat, x, N, K, bias))
    outs = ops.linear_forward_grouped(members, x, K)
    torch.cuda.synchronize()
    for y, ref in zip(outs, singles):
        assert torch.equal(y, ref)


def test_prefill_gemm_k_split_is_deterministic(ops, monkeypatch):
    """Long rows with few output tiles run K-split over gridDim.z (K >= 8192): the last CTA of a tile adds the partial tiles
    in split order, so reruns are bit-identical, the tile counters are left at zero (third run), and the result sits
    within tolerance of fp32 like the unsplit kernel's."""
    dev = "cuda"
    N, K, M, bits = 256, 8192, 40, 3
    codes, scale, zero = _synthetic(N, K, bits, seed=5)
    cg, sg, zg = torch.from_numpy(codes).to(dev), scale.to(dev), zero.to(dev)
    nat = ops.pack_native(bits, cg, sg, zg)
    W = (cg.float().reshape(N, K // G, G) * sg.float()[..., None] - (zg * sg).float()[..., None]).reshape(N, K)
    torch.manual_seed(3)
    x = torch.randn(M, K, device=dev).half()
    bias = torch.randn(N, device=dev).half()
    ref = (x.float() @ W.t() + bias.float()).cpu()
    ws = ops.gemm_workspace(M, K, bits, dev)
    ws.fill_(0xA5)                                             # a caller's workspace is uninitialised memory
    ys = [ops.gemm_tc(bits, nat, x, N, K, bias, workspace=ws).clone() for _ in range(3)]
    assert torch.equal(ys[0], ys[1]) and torch.equal(ys[1], ys[2])
    assert O.max_rel(ys[0].cpu(), ref) <= TOL
    monkeypatch.setenv("AMQB_TC_NO_SPLITK", "1")
    y1 = ops.gemm_tc(bits, nat, x, N, K, bias, workspace=ws)
    assert O.max_rel(y1.cpu(), ref) <= TOL
    assert O.max_rel(ys[0].cpu(), y1.float().cpu()) <= 1e-3


def test_prefill_cluster_variant(ops, monkeypatch):
    """The cluster-pair variant (X stages multicast to two CTAs) must give the same numbers as the default."""
    dev = "cuda"
    N, K, M, bits = 512, 1024, 200, 3
    codes, scale, zero = _synthetic(N, K, bits, seed=5)
    nat = ops.pack_native(bits, torch.from_numpy(codes).to(dev), scale.to(dev), zero.to(dev))
    x = torch.randn(M, K, device=dev).half()
    y0 = ops.gemm_tc(bits, nat, x, N, K)
    monkeypatch.setenv("AMQB_TC_CLUSTER", "1")
    y1 = ops.gemm_tc(bits, nat, x, N, K)
    torch.cuda.synchronize()
    assert torch.equal(y0, y1)


def test_error_conventions(ops):
    from amq_b200 import _lib
    L = _lib.lib()
    assert L.amqb_native_bytes(3, 100, 4096) == 0           # N % 32 != 0
    x = torch.zeros(17, 128, device="cuda", dtype=torch.float16)
    nat = torch.zeros(ops.native_bytes(3, 32, 128), dtype=torch.uint8, device="cuda")
    ws = ops.workspace(x.device, 64, 128, 16)
    y = torch.zeros(17, 32, device="cuda", dtype=torch.float16)
    rc = L.amqb_gemv_w3(_lib.ptr(nat), _lib.ptr(x), _lib.ptr(y), None, 17, 32, 128, _lib.ptr(ws), ws.numel(), None)
    assert rc == -1 and b"M must be 1..16" in L.amqb_last_error_string()
    rc = L.amqb_gemv_w3(_lib.ptr(nat), _lib.ptr(x), _lib.ptr(y), None, 1, 48, 128, _lib.ptr(ws), ws.numel(), None)
    assert rc == -2
    with pytest.raises(RuntimeError):
        ops.gemv(3, nat.cpu(), x[:1], 32, 128)              # CPU tensor: loud failure, no fallback
