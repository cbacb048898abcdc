"""Model-level parity: the captured decode step (C-ABI kernels, CUDA graph, PDL) against a plain
PyTorch fp32 decoder built from the SAME weights (dense weights recovered from the native buffers)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_step(model, tok, pos, kc, vc):
    """fp32 reference of one decode step with HF conventions (RMSNorm, rotate_half RoPE, GQA, SiLU MLP)."""
    from amq_b200 import ops
    S = model.shape
    D, Hq, Hkv = model.D, model.Hq, model.Hkv
    h = model.embed[tok].float()
    B = h.shape[0]

    def rms(x, g, eps):
        return g.float() * (x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps)).half().float()

    def dense(L, name):
        bits, nat, N, K = L[name]
        return ops.native_to_dense(bits, nat, N, K)

    inv = S.rope_theta ** (-torch.arange(0, D // 2, device=h.device).float() * 2 / D)
    ang = pos * inv
    cos = torch.cat([ang.cos(), ang.cos()]).half().float()
    sin = torch.cat([ang.sin(), ang.sin()]).half().float()

    def rope(x):          # x [B, n, D]
        x1, x2 = x[..., : D // 2], x[..., D // 2:]
        return (x * cos + torch.cat([-x2, x1], -1) * sin).half().float()

    for li, L in enumerate(model.layers):
        x = rms(h, L["norm1"], S.rms_eps).half().float()
        q = x @ dense(L, "self_attn.q_proj").t()
        k = x @ dense(L, "self_attn.k_proj").t()
        v = x @ dense(L, "self_attn.v_proj").t()
        if "qkv_bias" in L:
            b = L["qkv_bias"].float()
            q, k, v = q + b[: model.q_dim], k + b[model.q_dim: model.q_dim + model.kv_dim], v + b[model.q_dim + model.kv_dim:]
        q, k, v = q.half().float(), k.half().float(), v.half().float()
        q = rope(q.view(B, Hq, D))
        k = rope(k.view(B, Hkv, D))
        kc[li][:, :, pos] = k
        vc[li][:, :, pos] = v.view(B, Hkv, D)
        kk = kc[li][:, :, : pos + 1].repeat_interleave(Hq // Hkv, dim=1)
        vv = vc[li][:, :, : pos + 1].repeat_interleave(Hq // Hkv, dim=1)
        att = torch.softmax(torch.einsum("bhd,bhsd->bhs", q, kk) / D ** 0.5, -1)
        o = torch.einsum("bhs,bhsd->bhd", att, vv).reshape(B, Hq * D).half().float()
        h = (h + o @ dense(L, "self_attn.o_proj").t()).half().float()
        x = rms(h, L["norm2"], S.rms_eps).half().float()
        g = (x @ dense(L, "mlp.gate_proj").t()).half().float()
        u = (x @ dense(L, "mlp.up_proj").t()).half().float()
        a = (torch.nn.functional.silu(g).half().float() * u).half().float()
        h = (h + a @ dense(L, "mlp.down_proj").t()).half().float()
    x = rms(h, model.final_norm, S.rms_eps).half().float()
    return x @ model.lm_head.float().t()


@pytest.mark.parametrize("family", ["llama", "qwen2"])
@pytest.mark.parametrize("batch", [1, 3])
def test_decode_step_matches_torch_reference(family, batch):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from amq_b200.arch import ModelShape, LINEARS
    from amq_b200.model import QuantDecoder
    if family == "llama":
        shape = ModelShape("tiny-llama", 256, 512, 4, 4, 2, 512, head_dim=64)
    else:
        shape = ModelShape("tiny-qwen2", 256, 512, 4, 2, 2, 512, head_dim=64, rope_theta=1e6, rms_eps=1e-6, qkv_bias=True)
    rs = np.random.RandomState(0)
    arch = {n: rs.choice([2, 3, 4], size=2).tolist() for n in LINEARS}
    m = QuantDecoder(shape, arch, batch=batch, max_seq=32, seed=1)
    kc = [torch.zeros(batch, m.Hkv, 32, m.D, device=m.dev) for _ in m.layers]
    vc = [torch.zeros(batch, m.Hkv, 32, m.D, device=m.dev) for _ in m.layers]
    tok = torch.randint(0, shape.vocab, (batch,), device=m.dev)
    m.reset()
    m.tokens.copy_(tok)
    agree = 0
    steps = 6
    for pos in range(steps):
        cur = m.tokens.clone()
        ref = _ref_step(m, cur, pos, kc, vc)
        m.step()                                   # graph replay (captured on first call)
        torch.cuda.synchronize()
        got = m.logits.clone()
        rel = (got - ref).abs().max() / ref.abs().max()
        assert rel < 2e-2, (family, batch, pos, float(rel))
        agree += int((got.argmax(-1) == ref.argmax(-1)).sum())
        assert torch.equal(m.tokens, got.argmax(-1))          # argmax kernel + feedback
    assert agree >= int(0.8 * steps * batch)
    # eager (no graph, no PDL) path gives bit-identical logits to the graph path
    m2 = QuantDecoder(shape, arch, batch=batch, max_seq=32, seed=1)
    m2.reset(); m2.tokens.copy_(tok)
    m.reset(); m.tokens.copy_(tok)
    for pos in range(3):
        m.step(); m2.step_eager()
    torch.cuda.synchronize()
    assert torch.equal(m.logits, m2.logits)


def test_step_host_matches_device_loop():
    """The host-facing step (pinned ids in / out as memcpy nodes of the captured graph) generates the same tokens
    as the device-resident loop."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from amq_b200.arch import ModelShape, LINEARS
    from amq_b200.model import QuantDecoder
    shape = ModelShape("tiny-llama", 256, 512, 4, 4, 2, 512, head_dim=64)
    arch = {n: [3, 4] for n in LINEARS}
    B = 2
    m = QuantDecoder(shape, arch, batch=B, max_seq=32, seed=3)
    start = torch.tensor([5, 17], dtype=torch.int64)
    m.reset(); m.tokens.copy_(start.to(m.dev))
    want = []
    for _ in range(5):
        m.step()
        want.append(m.tokens.cpu().clone())
    host_in, host_out = start.clone().pin_memory(), torch.zeros(B, dtype=torch.int64).pin_memory()
    m.reset()
    got = []
    for _ in range(5):
        m.step_host(host_in, host_out)          # returns after the D2H copy has landed
        got.append(host_out.clone())
        host_in.copy_(host_out)
    assert all(torch.equal(a, b) for a, b in zip(want, got))
    with pytest.raises(ValueError):
        m.step_host(start.clone(), host_out)    # not pinned


@pytest.mark.gpu
@pytest.mark.parametrize("family,batch,prompt", [("llama", 2, 9), ("llama", 1, 42), ("qwen2", 2, 26), ("gqa128", 1, 70)])
def test_prefill_matches_token_by_token(family, batch, prompt):
    """QuantDecoder.prefill (one pass over the weights for all prompt rows: tcgen05 GEMM above 16 rows, skinny decode
    kernel below, row kernels of csrc/prefill_glue.cu in between) leaves the same K/V cache as the same prompt
    consumed token by token through the decode step, and the next decode step gives the same logits, to the fp16
    rounding of the linears.  Also in two chunks (pos0 > 0)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from amq_b200.arch import ModelShape, LINEARS
    from amq_b200.model import QuantDecoder
    if family == "llama":
        shape = ModelShape("tiny-llama", 256, 512, 4, 4, 3, 512, head_dim=64)
    elif family == "qwen2":
        shape = ModelShape("tiny-qwen2", 256, 512, 4, 2, 2, 512, head_dim=64, rope_theta=1e6, rms_eps=1e-6, qkv_bias=True)
    else:
        shape = ModelShape("tiny-gqa", 512, 1408, 4, 2, 2, 512, head_dim=128)
    rs = np.random.RandomState(2)
    arch = {n: rs.choice([2, 3, 4], size=shape.n_block).tolist() for n in LINEARS}
    S = 96
    m = QuantDecoder(shape, arch, batch=batch, max_seq=S, seed=4)
    ids = torch.randint(0, shape.vocab, (batch, prompt), device=m.dev)
    P = prompt - 1

    def last_step():
        m.tokens.copy_(ids[:, P]); m.step(); torch.cuda.synchronize()
        return m.logits.clone()

    def caches():
        return [(L["k_cache"][:, :, :P].float().clone(), L["v_cache"][:, :, :P].float().clone()) for L in m.layers]

    def wipe():
        for L in m.layers:
            L["k_cache"].zero_(); L["v_cache"].zero_()
        m.reset()

    wipe()
    for t in range(P):
        m.tokens.copy_(ids[:, t]); m.step()
    want_c, want_l = caches(), last_step()

    def compare(tag):
        assert int(m.pos.item()) == P, tag
        for li, ((k0, v0), (k1, v1)) in enumerate(zip(want_c, caches())):
            assert (k0 - k1).abs().max() / k0.abs().max() < 1e-2, (tag, family, li, "k")
            assert (v0 - v1).abs().max() / v0.abs().max() < 1e-2, (tag, family, li, "v")
        got_l = last_step()
        assert (got_l - want_l).abs().max() / want_l.abs().max() < 2e-2, (tag, family)

    wipe()
    m.prefill(ids[:, :P])
    compare("one pass")
    wipe()
    cut = P // 3 + 1                                  # second chunk attends to cache rows written by the first
    m.prefill(ids[:, :cut]); m.prefill(ids[:, cut:P])
    compare("two chunks")
    wipe()
    m.prefill(ids[:, :P], use_graph=False)            # plain launches (the default replays a captured graph)
    compare("eager")
    wipe()
    m.prefill(ids[:, :P])                             # replay of the graph captured above, after the cache was wiped
    compare("graph replay")
    # generate(): prefill and token-by-token prompts give the same first token when its margin is clear
    a = m.generate(ids, 3, prefill=True)
    b = m.generate(ids, 3, prefill=False)
    top2 = want_l.topk(2, dim=-1).values
    clear = (top2[:, 0] - top2[:, 1]) > 4e-2 * want_l.abs().max()
    assert torch.equal(a[clear, 0], b[clear, 0])


@pytest.mark.gpu
def test_speed_benchmark_protocol(tmp_path, monkeypatch):
    """benchmark_speed (the reference's protocol, amq/utils/speed.py:130-255) in all four modes on a tiny decoder, and
    the amq_speed_benchmark.py command line on one block of Llama-2-7B with a searched-arch file."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import json
    import os
    import sys
    from amq_b200.arch import MODELS, ModelShape, LINEARS, make_stats_file
    from amq_b200.model import QuantDecoder
    from amq_b200.utils.speed import benchmark_speed
    shape = ModelShape("tiny-llama", 256, 512, 4, 4, 2, 512, head_dim=64)
    arch = {n: [2, 4] for n in LINEARS}
    m = QuantDecoder(shape, arch, batch=2, max_seq=64, seed=5)
    for mode, it in (("TPS", 2), ("GeMV", 1), ("GeMM", 3), ("TTFT", 3)):
        d = benchmark_speed(m, None, iteration=it, sizes=(2, 20, 8), mode=mode, get_peak_memory=(mode == "TPS"))
        assert set(d) == ({mode.lower(), "peak_memory"} if mode == "TPS" else {mode.lower()})
        assert d[mode.lower()]["2.20.8"] > 0
    with pytest.raises(ValueError):
        benchmark_speed(m, None, sizes=(1, 20, 8), mode="TPS")          # batch mismatch
    with pytest.raises(ValueError):
        benchmark_speed(m, None, sizes=(2, 60, 8), mode="TPS")          # beyond the static KV cache
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import amq_speed_benchmark as cli
    stats = str(tmp_path / "iter_1.stats")
    make_stats_file(stats, MODELS["Llama-2-7b-hf"], 3.0, n=4)
    monkeypatch.chdir(tmp_path)
    res = cli.main(["--model_name", "Llama-2-7b-hf", "--n_block", "1", "--target_bits", "3", "--arch_path", stats, "--tps",
                    "--gemv", "--memory", "--seq_length", "24", "--gen_length", "8", "--file_name", "out.json"])
    assert res["3.0bit"]["tps"]["1.24.8"] > 0 and res["3.0bit"]["gemv"]["1.24.8"] > 0 and res["3.0bit"]["memory"] > 0.5
    saved = json.load(open(tmp_path / "benchmark" / "outputs" / "out.json"))
    assert saved["args"]["target_bits"] == 3.0 and "3.0bit" in saved


@pytest.mark.gpu
@pytest.mark.parametrize("M,H", [(1, 256), (19, 512), (63, 4096), (40, 8192)])
def test_add_rmsnorm_rows_equals_the_two_calls(M, H):
    """amqb_add_rmsnorm_rows (residual add + RMSNorm of the updated rows in one launch, what follows every o_proj /
    down_proj of the prompt pass) is bit-identical to amqb_add_rows followed by amqb_rmsnorm_rows."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ctypes
    from amq_b200._lib import check, cur_stream, lib, ptr
    dev = "cuda:0"
    torch.cuda.set_device(0)
    g = torch.Generator(device=dev).manual_seed(M + H)
    L, st = lib(), cur_stream()
    h0 = torch.randn(M, H, device=dev, generator=g).half()
    y = (0.3 * torch.randn(M, H, device=dev, generator=g)).half()
    gamma = (1 + 0.1 * torch.randn(H, device=dev, generator=g)).half()
    eps = ctypes.c_float(1e-5)
    ha, xa = h0.clone(), torch.empty_like(h0)
    check(L.amqb_add_rows(ptr(ha), ptr(y), M, H, st), "add_rows")
    check(L.amqb_rmsnorm_rows(ptr(ha), ptr(gamma), eps, ptr(xa), M, H, st), "rmsnorm_rows")
    hb, xb = h0.clone(), torch.empty_like(h0)
    check(L.amqb_add_rmsnorm_rows(ptr(hb), ptr(y), ptr(gamma), eps, ptr(xb), M, H, st), "add_rmsnorm_rows")
    torch.cuda.synchronize()
    assert torch.equal(ha, hb) and torch.equal(xa, xb)
    assert torch.equal(ha, (h0.float() + y.float()).half())
    assert L.amqb_add_rmsnorm_rows(ptr(hb), ptr(y), ptr(gamma), eps, ptr(hb), M, H, st) != 0      # out must not alias h


@pytest.mark.gpu
@pytest.mark.parametrize("D,Hq,Hkv,B,T,pos0", [(64, 4, 2, 2, 19, 0), (128, 8, 2, 1, 45, 7), (128, 4, 4, 3, 8, 33)])
def test_prefill_row_kernels_against_torch(D, Hq, Hkv, B, T, pos0):
    """Each kernel of csrc/prefill_glue.cu against a plain PyTorch fp32 statement of the same op: row RMSNorm, SiLU*up,
    residual add, RoPE (q in place, k into the cache), V append, causal attention over [earlier context | this block]."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ctypes
    from amq_b200._lib import check, cur_stream, lib, ptr
    dev, S, theta = "cuda:0", 64, 10000.0
    torch.cuda.set_device(0)
    g = torch.Generator(device=dev).manual_seed(D + T)
    M, H = B * T, Hq * D
    L, st = lib(), cur_stream()

    def rnd(*shape):
        return torch.randn(*shape, device=dev, generator=g).half()

    def close(got, ref, tol):
        return float((got.float() - ref).abs().max()) <= tol * float(ref.abs().max())

    x, gamma = rnd(M, H), (1 + 0.1 * torch.randn(H, device=dev, generator=g)).half()
    out = torch.empty_like(x)
    check(L.amqb_rmsnorm_rows(ptr(x), ptr(gamma), ctypes.c_float(1e-5), ptr(out), M, H, st), "rmsnorm_rows")
    xf = x.float()
    ref = gamma.float() * (xf * torch.rsqrt((xf * xf).mean(-1, keepdim=True) + 1e-5)).half().float()
    assert close(out, ref, 1e-3)
    gate, up = rnd(M, H), rnd(M, H)
    check(L.amqb_silu_mul_rows(ptr(gate), ptr(up), ptr(out), M, H, st), "silu_mul_rows")
    assert close(out, torch.nn.functional.silu(gate.float()).half().float() * up.float(), 1e-3)
    h0, y = rnd(M, H), rnd(M, H)
    h = h0.clone()
    check(L.amqb_add_rows(ptr(h), ptr(y), M, H, st), "add_rows")
    assert torch.equal(h, (h0.float() + y.float()).half())

    rope = torch.empty(S, D // 2, 2, dtype=torch.float32, device=dev)
    check(L.amqb_rope_table(ptr(rope), S, D, ctypes.c_float(theta), st), "rope_table")
    q, k, v = rnd(M, Hq * D), rnd(M, Hkv * D), rnd(M, Hkv * D)
    kc = torch.zeros(B, Hkv, S, D, device=dev, dtype=torch.float16)
    vc = torch.zeros(B, Hkv, S, D, device=dev, dtype=torch.float16)
    kc[:, :, :pos0], vc[:, :, :pos0] = rnd(B, Hkv, pos0, D), rnd(B, Hkv, pos0, D)        # earlier context
    q0, kc0, vc0 = q.clone(), kc.clone(), vc.clone()
    att = torch.empty(M, Hq * D, device=dev, dtype=torch.float16)
    check(L.amqb_attn_prefill(ptr(q), ptr(k), ptr(v), ptr(kc), ptr(vc), ptr(att), pos0, T, B, Hq, Hkv, D, S, ptr(rope), st),
          "attn_prefill")
    inv = theta ** (-torch.arange(0, D // 2, device=dev).float() * 2 / D)
    ang = torch.arange(pos0, pos0 + T, device=dev).float()[:, None] * inv[None]
    cos = torch.cat([ang.cos(), ang.cos()], -1).half().float()[None, :, None, :]
    sin = torch.cat([ang.sin(), ang.sin()], -1).half().float()[None, :, None, :]

    def rot(t):                   # [B, T, heads, D], HF rotate_half
        t1, t2 = t[..., : D // 2], t[..., D // 2:]
        return (t * cos + torch.cat([-t2, t1], -1) * sin).half().float()

    qr, kr = rot(q0.float().view(B, T, Hq, D)), rot(k.float().view(B, T, Hkv, D))
    assert close(q.view(B, T, Hq, D), qr, 2e-3)
    Kf, Vf = kc0.float(), vc0.float()
    Kf[:, :, pos0:pos0 + T] = kr.permute(0, 2, 1, 3)
    Vf[:, :, pos0:pos0 + T] = v.float().view(B, T, Hkv, D).permute(0, 2, 1, 3)
    assert close(kc, Kf, 2e-3) and torch.equal(vc.float(), Vf)
    assert torch.equal(kc[:, :, pos0 + T:], kc0[:, :, pos0 + T:])                          # nothing written past the block
    # attention reference from the kernel's own (fp16) rotated q / cache contents: isolates the softmax . V part
    Kr = kc.float()[:, :, : pos0 + T].repeat_interleave(Hq // Hkv, dim=1)
    Vr = vc.float()[:, :, : pos0 + T].repeat_interleave(Hq // Hkv, dim=1)
    sc = torch.einsum("bthd,bhsd->bhts", q.float().view(B, T, Hq, D), Kr) / D ** 0.5
    vis = torch.arange(pos0 + T, device=dev)[None, :] <= (pos0 + torch.arange(T, device=dev))[:, None]
    sc = sc.masked_fill(~vis[None, None], float("-inf"))
    o = torch.einsum("bhts,bhsd->bthd", torch.softmax(sc, -1), Vr).reshape(M, Hq * D)
    assert close(att, o, 2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("D,Hq,Hkv,B,pos,splits", [(128, 8, 2, 1, 300, 4), (64, 4, 4, 2, 517, 3), (128, 4, 2, 1, 40, 4),
                                                    (128, 32, 32, 1, 256, 4)])
def test_attn_decode_split_matches_single_cta(D, Hq, Hkv, B, pos, splits):
    """amqb_attn_decode_split (cached positions of a head shared by several CTAs beyond split_min_pos, last CTA merges)
    against the single-CTA kernel and a PyTorch fp32 attention over the same cache; below the threshold it is
    bit-identical; replays leave the arrival counters clean."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ctypes
    from amq_b200._lib import check, cur_stream, lib, ptr
    dev, S, theta, min_pos = "cuda:0", 640, 10000.0, 256
    torch.cuda.set_device(0)
    g = torch.Generator(device=dev).manual_seed(pos)
    L, st = lib(), cur_stream()

    def rnd(*shape):
        return torch.randn(*shape, device=dev, generator=g).half()

    rope = torch.empty(S, D // 2, 2, dtype=torch.float32, device=dev)
    check(L.amqb_rope_table(ptr(rope), S, D, ctypes.c_float(theta), st), "rope_table")
    qkv = rnd(B, (Hq + 2 * Hkv) * D)
    kc0 = torch.zeros(B, Hkv, S, D, device=dev, dtype=torch.float16)
    vc0 = torch.zeros(B, Hkv, S, D, device=dev, dtype=torch.float16)
    kc0[:, :, :pos], vc0[:, :, :pos] = rnd(B, Hkv, pos, D), rnd(B, Hkv, pos, D)
    pos_dev = torch.tensor([pos], dtype=torch.int32, device=dev)
    kc1, vc1, out1 = kc0.clone(), vc0.clone(), torch.empty(B, Hq * D, device=dev, dtype=torch.float16)
    check(L.amqb_attn_decode(ptr(qkv), ptr(kc1), ptr(vc1), ptr(out1), ptr(pos_dev), B, Hq, Hkv, D, S, ctypes.c_float(theta),
                             ptr(rope), st), "attn_decode")
    L.amqb_attn_split_workspace_bytes.restype = ctypes.c_size_t
    ws = torch.zeros(int(L.amqb_attn_split_workspace_bytes(B, Hq, D, splits)), dtype=torch.uint8, device=dev)
    outs = []
    for _ in range(3):                                   # same workspace: the counters must come back to zero
        kc2, vc2, out2 = kc0.clone(), vc0.clone(), torch.empty(B, Hq * D, device=dev, dtype=torch.float16)
        check(L.amqb_attn_decode_split(ptr(qkv), ptr(kc2), ptr(vc2), ptr(out2), ptr(pos_dev), B, Hq, Hkv, D, S,
                                       ctypes.c_float(theta), ptr(rope), splits, min_pos, ptr(ws), ctypes.c_size_t(ws.numel()), st),
              "attn_decode_split")
        outs.append(out2)
        assert torch.equal(kc1, kc2) and torch.equal(vc1, vc2)
    torch.cuda.synchronize()
    assert int(ws[: 4 * B * Hq].view(torch.int32).abs().sum()) == 0
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    if pos < min_pos:
        assert torch.equal(out1, outs[0])
    # fp32 reference from the cache the kernel left (rotated k of this step at row pos) and the rotated q
    inv = theta ** (-torch.arange(0, D // 2, device=dev).float() * 2 / D)
    ang = pos * inv
    cos = torch.cat([ang.cos(), ang.cos()]).half().float()
    sin = torch.cat([ang.sin(), ang.sin()]).half().float()
    q = qkv[:, : Hq * D].float().view(B, Hq, D)
    q = (q * cos + torch.cat([-q[..., D // 2:], q[..., : D // 2]], -1) * sin).half().float()
    K = kc1.float()[:, :, : pos + 1].repeat_interleave(Hq // Hkv, dim=1)
    V = vc1.float()[:, :, : pos + 1].repeat_interleave(Hq // Hkv, dim=1)
    att = torch.softmax(torch.einsum("bhd,bhsd->bhs", q, K) / D ** 0.5, -1)
    ref = torch.einsum("bhs,bhsd->bhd", att, V).reshape(B, Hq * D)
    for o in (out1, outs[0]):
        assert float((o.float() - ref).abs().max()) <= 2e-3 * float(ref.abs().max())


@pytest.mark.gpu
def test_decode_step_with_split_attention_in_graph(monkeypatch):
    """The captured decode step with the split attention path forced on from position 8: logits against the PyTorch
    fp32 decoder across the threshold (graph replays reuse one workspace for every layer)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from amq_b200.arch import ModelShape, LINEARS
    from amq_b200.model import QuantDecoder
    monkeypatch.setenv("AMQB_ATTN_SPLIT_MIN_POS", "8")
    shape = ModelShape("tiny-gqa", 512, 1408, 4, 2, 2, 512, head_dim=128)
    arch = {n: [3, 4] for n in LINEARS}
    m = QuantDecoder(shape, arch, batch=2, max_seq=32, seed=7)
    assert m.attn_splits == 4 and m.attn_split_min_pos == 8
    kc = [torch.zeros(2, m.Hkv, 32, m.D, device=m.dev) for _ in m.layers]
    vc = [torch.zeros(2, m.Hkv, 32, m.D, device=m.dev) for _ in m.layers]
    m.reset(); m.tokens.copy_(torch.tensor([3, 11], device=m.dev))
    for pos in range(14):
        ref = _ref_step(m, m.tokens.clone(), pos, kc, vc)
        m.step()
        torch.cuda.synchronize()
        assert (m.logits - ref).abs().max() / ref.abs().max() < 2e-2, pos


@pytest.mark.parametrize("batch", [1, 4, 16])
def test_full_size_llama7b_layer_matches_torch_reference(batch):
    """One REAL Llama-2-7B decoder layer (hidden 4096, inter 11008, 32 heads; amq/configs/llama.json:2-27) with mixed
    bit widths, in-model, against the fp32 PyTorch decoder from the same weights: exercises the 148-CTA grouped launches,
    the K = 11008 chunked x' of down_proj, the q|k|v / gate|up x' variant sharing and, at batch 4 / 16, the multi-row
    decode kernels inside the captured step."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dataclasses
    from amq_b200.arch import LINEARS, MODELS
    from amq_b200.model import QuantDecoder
    shape = dataclasses.replace(MODELS["Llama-2-7b-hf"], n_block=1, vocab=4096)
    arch = {n: [b] for n, b in zip(LINEARS, [2, 3, 4, 3, 4, 2, 3])}
    m = QuantDecoder(shape, arch, batch=batch, max_seq=16, seed=4)
    kc = [torch.zeros(batch, m.Hkv, 16, m.D, device=m.dev) for _ in m.layers]
    vc = [torch.zeros(batch, m.Hkv, 16, m.D, device=m.dev) for _ in m.layers]
    tok = torch.randint(0, shape.vocab, (batch,), device=m.dev)
    m.reset(); m.tokens.copy_(tok)
    for pos in range(4):
        cur = m.tokens.clone()
        ref = _ref_step(m, cur, pos, kc, vc)
        if pos == 0: m.step_eager()
        else: m.step()
        torch.cuda.synchronize()
        rel = float((m.logits - ref).abs().max() / ref.abs().max())
        assert rel < 2e-2, (batch, pos, rel)


def test_full_size_llama70b_layer_and_tp8_shards():
    """One real Llama-2-70B layer (hidden 8192, inter 28672, 64 / 8 heads; llama.json:56-81): unsharded against the fp32
    PyTorch decoder, then as 8 tensor-parallel shards (k / v projections of 128 rows, o_proj with K = 1024, down_proj with
    K = 3584: the cluster split-K and short-K launch geometries of config 5) against the unsharded step."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dataclasses
    from amq_b200 import tp
    from amq_b200.arch import LINEARS, MODELS
    from amq_b200.model import QuantDecoder
    shape = dataclasses.replace(MODELS["Llama-2-70b-hf"], n_block=1, vocab=2048)
    arch = {n: [b] for n, b in zip(LINEARS, [3, 2, 4, 3, 2, 4, 3])}
    full = QuantDecoder(shape, arch, batch=1, max_seq=16, seed=6)
    kc = [torch.zeros(1, full.Hkv, 16, full.D, device=full.dev)]
    vc = [torch.zeros(1, full.Hkv, 16, full.D, device=full.dev)]
    grp = tp.LocalTPGroup(full, 8)
    tok = torch.tensor([7], device=full.dev)
    full.reset(); full.tokens.copy_(tok)
    grp.set_tokens(tok)
    for pos in range(4):
        cur = full.tokens.clone()
        ref = _ref_step(full, cur, pos, kc, vc)
        for r in grp.ranks:
            r.tokens.copy_(cur)
        torch.cuda.synchronize()
        if pos == 0: full.step_eager(); grp.step_eager()
        else: full.step(); grp.step()
        torch.cuda.synchronize()
        rel = float((full.logits - ref).abs().max() / ref.abs().max())
        assert rel < 2e-2, ("full", pos, rel)
        rel = float((grp.ranks[0].logits - full.logits).abs().max() / full.logits.abs().max())
        assert rel < 2e-2, ("tp8", pos, rel)
        assert all(torch.equal(r.logits, grp.ranks[0].logits) for r in grp.ranks)
    assert grp.timeouts() == 0
