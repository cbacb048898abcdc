"""Generate tests/golden/*.npz from the UNMODIFIED reference and pin the oracle.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

It imports the reference's vendored HQQ (amq/kernel/hqq) with a stub
``termcolor`` module (its only missing import), runs the reference's
Quantizer / BitPack / GPTQLinear / pack_intweight on seeded inputs, asserts
that every oracle function in oracle/amq_oracle.py reproduces the reference
output (bit-exact for integer work, exact equality for the fp16 torch path),
and stores inputs + reference outputs as small fixtures.  The fixtures are
committed; the reference is not needed to run the tests.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/amq/kernel/hqq"
sys.path.insert(0, ROOT)


def import_reference():
    if "termcolor" not in sys.modules:
        tc = types.ModuleType("termcolor")
        tc.colored = lambda s, *a, **k: s
        sys.modules["termcolor"] = tc
    sys.path.insert(0, REF)
    from hqq.core.bitpack import BitPack
    from hqq.core.quantize import Quantizer, BaseQuantizeConfig
    from hqq.backends.autogptq import GPTQLinear
    from hqq.backends.ft import pack_intweight
    return BitPack, Quantizer, BaseQuantizeConfig, GPTQLinear, pack_intweight


def fp16_solver_fixtures(out_dir, report):
    """The reference's CUDA branch runs the solver in fp16 (optimize.py:231).  No GPU here, so its own step function
    (optimize_weights_proximal_legacy_step, unmodified) is driven on fp16 CPU tensors through the same loop as
    optimize_weights_proximal_legacy (:234-247) and the final rounding of :254; the oracle's solver_dtype=fp16 mode must
    agree exactly.  This pins the fp16 MODE's op sequence; the reference's CUDA kernels may still round a pow or a mean
    differently from torch's CPU fp16 kernels in rare last-bit cases."""
    from oracle import amq_oracle as O
    from hqq.core.optimize import optimize_weights_proximal_legacy_step as step
    G = 128
    for nbits in (2, 3, 4):
        torch.manual_seed(2000 + nbits)
        N, K = 256, 512
        W = (torch.randn(N, K) * 0.02).half()
        Wf = W.float().reshape(-1, G)
        mn, mx = Wf.min(1, keepdim=True)[0], Wf.max(1, keepdim=True)[0]
        max_v = round(2 ** nbits - 1)
        scale = (max_v / (mx - mn))
        scale = torch.where((mx - mn).abs() <= 1e-4, torch.full_like(scale, 1.0), scale).clamp(max=2e4)
        zero = -mn * scale
        if nbits == 4:
            zero = torch.round(zero)
        W_f, sc, ze = Wf.half(), scale.half(), zero.half()
        best = torch.tensor(torch.inf, dtype=torch.float32)
        n_it = 0
        for _ in range(20):
            n_it += 1
            W_r, W_q, ze, sc = step(W_f, sc, ze, [0, max_v], 1e1, 0.7, 1)
            cur = torch.abs(W_f - W_r).mean().float()
            if cur < best:
                best = cur
            else:
                break
        W_q = torch.round(Wf * sc + ze).clamp_(0, max_v)
        codes, o_scale, o_zero, o_it = O.hqq_quantize(W, nbits, G, solver_dtype=torch.float16)
        assert o_it == n_it, ("fp16 solver iterations", nbits, o_it, n_it)
        assert np.array_equal(codes, W_q.to(torch.uint8).numpy()), ("fp16 solver codes", nbits)
        assert torch.equal(o_zero, ze) and torch.equal(o_scale, 1.0 / sc), ("fp16 solver meta", nbits)
        np.savez_compressed(os.path.join(out_dir, f"hqq_fp16solver_{nbits}bit.npz"), W=W.numpy(), codes=codes,
                            scale=o_scale.float().numpy(), zero=o_zero.float().numpy(), solver_iters=np.int32(n_it))
    report.append("fp16-solver fixtures (reference step function on fp16 CPU tensors) ok")


def _tiny_llama(seed):
    import transformers
    cfg = transformers.LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=2, vocab_size=512, max_position_embeddings=128, rms_norm_eps=1e-5,
                                   tie_word_embeddings=False)
    torch.manual_seed(seed)
    model = transformers.LlamaForCausalLM(cfg).half().eval()
    for p in model.parameters():                      # N(0, 0.02) linears like the survey's parity recipe
        p.requires_grad = False
    return model


def proxy_checkpoint_fixture(out_dir, report):
    """A tiny Llama quantized and SAVED BY THE REFERENCE (AutoHQQHFModel.quantize_model + save_quantized,
    models/base.py:267-434) -> tests/golden/qmodel_tiny_llama_3bit/{config.json, qmodel.pt}, plus the reference model's own
    logits on a fixed prompt (CPU).  `accelerate` (only used by the reference's create_model) is stubbed."""
    if "accelerate" not in sys.modules:
        acc = types.ModuleType("accelerate")
        import contextlib
        acc.init_empty_weights = contextlib.nullcontext
        sys.modules["accelerate"] = acc
    from hqq.models.hf.base import AutoHQQHFModel
    from hqq.core.quantize import BaseQuantizeConfig, HQQLinear
    model = _tiny_llama(7)
    AutoHQQHFModel.quantize_model(model, BaseQuantizeConfig(nbits=3, group_size=128), compute_dtype=torch.float16, device="cpu")
    save_dir = os.path.join(out_dir, "qmodel_tiny_llama_3bit")
    AutoHQQHFModel.save_quantized(model, save_dir)
    ids = torch.tensor([[3, 17, 256, 99, 5, 480, 42, 7]])
    with torch.no_grad():
        logits = model(ids).logits.float()
    torch.save({"input_ids": ids, "logits": logits}, os.path.join(save_dir, "reference_logits.pt"))
    n_q = sum(type(m) is HQQLinear for m in model.modules())
    report.append(f"proxy checkpoint written by the reference: {n_q} HQQLinear layers, logits {tuple(logits.shape)}")


def hf_dropin_fixture(out_dir, report):
    """The survey's model-level oracle: a tiny HF Llama whose 14 linears are the REFERENCE's GPTQLinear
    modules (mixed 2/3/4 bits: HQQ quantize -> dequantize -> GPTQLinear.pack, kernel_switch_threshold = 0 so that forward
    is the reference's own torch dequant + matmul, autogptq.py:245-283), assembled by setattr like
    amq/amq_speed_benchmark.py:231-251, run on CPU: logits of a prompt and greedy tokens.  The fixture carries the fp16
    model (state dict of the non-quantized parts), every linear's reference buffers and bit width, and the outputs."""
    from hqq.core.quantize import Quantizer, BaseQuantizeConfig
    from hqq.backends.autogptq import GPTQLinear
    model = _tiny_llama(11)
    rs = np.random.RandomState(5)
    names = ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
             "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"]
    arch = {n: rs.choice([2, 3, 4], size=2).tolist() for n in names}
    linears = {}
    for li, layer in enumerate(model.model.layers):
        for name in names:
            mod, lin = name.split(".")
            src = getattr(getattr(layer, mod), lin)
            bits = int(arch[name][li])
            W = src.weight.data
            cfg = BaseQuantizeConfig(nbits=bits, group_size=128)["weight_quant_params"]
            W_q, meta = Quantizer.quantize(W, device="cpu", compute_dtype=torch.float16, **cfg)
            meta16 = dict(meta, scale=meta["scale"].half(), zero=meta["zero"].half(), compute_dtype=torch.float16)
            W_deq = Quantizer.dequantize(W_q, meta16)
            N, K = W.shape
            g = GPTQLinear(bits, 128, K, N, bias=False, kernel_switch_threshold=0)
            g.pack(W_deq, meta16["scale"].reshape(N, -1), meta16["zero"].reshape(N, -1))
            delattr(getattr(layer, mod), lin)
            setattr(getattr(layer, mod), lin, g)
            linears[f"{li}.{name}"] = {"bits": bits, "qweight": g.qweight.clone(), "scales": g.scales.clone(), "zeros": g.zeros.clone()}
    ids = torch.tensor([[3, 17, 256, 99, 5, 480, 42, 7]])
    with torch.no_grad():
        logits = model(ids).logits.float()
        gen = model.generate(ids, max_new_tokens=8, do_sample=False)
    rest = {k: v.clone() for k, v in model.state_dict().items() if not any(t in k for t in ("qweight", "scales", "zeros"))}
    torch.save({"config": model.config.to_dict(), "arch": arch, "linears": linears, "rest": rest, "input_ids": ids,
                "logits": logits, "generated": gen}, os.path.join(out_dir, "hf_dropin_tiny_llama.pt"))
    report.append(f"HF drop-in fixture: reference GPTQLinear modules in a tiny Llama, greedy tokens {gen[0, ids.shape[1]:].tolist()}")


def hqqlinear_state_fixtures(out_dir, report):
    """A state dict WRITTEN BY THE REFERENCE's HQQLinear (quantize.py:643-682), in both forms the reference produces:
    encoded (the default: every non-tensor entry as a tensor, core/utils.py:37-69) and plain (what
    BaseHQQModel.serialize_weights stores in qmodel.pt, models/base.py:405-422), plus the reference's own forward on it."""
    import torch.nn as nn
    from hqq.core.quantize import HQQLinear, BaseQuantizeConfig
    for nbits in (2, 3, 4):
        torch.manual_seed(50 + nbits)
        lin = nn.Linear(256, 64, bias=(nbits != 2)).half()
        lin.weight.data = (torch.randn(64, 256) * 0.02).half()
        layer = HQQLinear(lin, BaseQuantizeConfig(nbits=nbits, group_size=128), compute_dtype=torch.float16, device="cpu")
        enc = {k: (v.data.clone() if isinstance(v, torch.Tensor) else v) for k, v in layer.state_dict().items()}
        layer.encoded_state_dict = False
        plain = {k: (v.data.clone() if isinstance(v, torch.Tensor) else v) for k, v in layer.state_dict().items()}
        torch.manual_seed(3)
        x = torch.randn(5, 256).half()
        y = layer.forward_pytorch(x)
        y32 = x.float() @ layer.dequantize().float().t() + (layer.bias.float() if layer.bias is not None else 0.0)
        torch.save({"encoded": enc, "plain": plain, "x": x, "y_ref_fp16": y, "y_fp32": y32, "W_deq": layer.dequantize()},
                   os.path.join(out_dir, f"hqqlinear_state_{nbits}bit.pt"))
    report.append("HQQLinear state dicts (encoded + plain) written by the reference")


def main():
    from oracle import amq_oracle as O
    BitPack, Quantizer, BaseQuantizeConfig, GPTQLinear, pack_intweight = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    G = 128
    report = []

    # ---- 1. BitPack round trips (the reference's tests/test_bitpack.py property)
    rng = np.random.RandomState(42)
    for nbits, pack, unpack in [(4, BitPack.pack_4bit_u8, BitPack.unpack_4bit_u8),
                                (2, BitPack.pack_2bit_u8, BitPack.unpack_2bit_u8),
                                (3, BitPack.pack_3bit_32, BitPack.unpack_3bit_32)]:
        for R in (40, 64, 130):
            if nbits != 3 and R % {4: 2, 2: 4}[nbits]:
                continue
            codes = rng.randint(0, 2 ** nbits, size=(R, G)).astype(np.uint8)
            ref_p = pack(torch.from_numpy(codes)).numpy()
            assert np.array_equal(O.hqq_pack(codes, nbits), ref_p), ("hqq_pack", nbits, R)
            ref_u = unpack(torch.from_numpy(ref_p)).numpy()
            assert np.array_equal(O.hqq_unpack(ref_p, nbits), ref_u), ("hqq_unpack", nbits, R)
            assert np.array_equal(O.hqq_unpack(ref_p, nbits, rows=R), codes)
            if R == 130 or (R == 64 and nbits != 3):
                np.savez_compressed(os.path.join(out_dir, f"bitpack_{nbits}bit_R{R}.npz"),
                                    codes=codes, packed=ref_p)
    report.append("bitpack ok")

    # ---- 2. Quantizer.quantize / dequantize -> GPTQLinear.pack -> forward (torch path)
    for (N, K) in [(64, 256), (256, 512)]:
        for nbits in (2, 3, 4):
            torch.manual_seed(1000 * nbits + N)
            W = (torch.randn(N, K) * 0.02).half()
            cfg = BaseQuantizeConfig(nbits=nbits, group_size=G)["weight_quant_params"]
            W_q, meta = Quantizer.quantize(W, device="cpu", compute_dtype=torch.float16, **cfg)
            meta["compute_dtype"] = torch.float16
            # the reference moves meta to fp16 on .cuda() (quantize.py:202-217)
            scale16 = meta["scale"].half()
            zero16 = meta["zero"].half()
            meta16 = dict(meta, scale=scale16, zero=zero16)
            W_deq = Quantizer.dequantize(W_q, meta16)

            # oracle: quantize
            codes, o_scale, o_zero, n_it = O.hqq_quantize(W, nbits, G)
            ref_codes = O.hqq_unpack(W_q.numpy(), nbits, rows=N * K // G)
            assert np.array_equal(codes, ref_codes), ("quantize codes", nbits)
            assert torch.equal(o_scale, meta["scale"]) and torch.equal(o_zero, meta["zero"]), ("meta", nbits)
            assert np.array_equal(O.hqq_pack(codes, nbits), W_q.numpy())
            o_deq = O.hqq_dequantize(codes, scale16, zero16, (N, K))
            assert torch.equal(o_deq, W_deq), ("dequant", nbits)

            # GPTQ pack
            layer = GPTQLinear(nbits, G, K, N, bias=False, kernel_switch_threshold=0)
            s2 = scale16.reshape(N, -1)
            z2 = zero16.reshape(N, -1)
            layer.pack(W_deq, s2, z2)
            q_codes = O.gptq_codes_from_weight(W_deq, s2, z2, G)
            assert np.array_equal(q_codes.astype(np.uint8), codes.reshape(N, K)), ("codes survive dequant", nbits)
            assert np.array_equal(O.gptq_pack_codes(q_codes, nbits), layer.qweight.numpy()), ("gptq pack", nbits)
            assert np.array_equal(O.gptq_unpack(layer.qweight.numpy(), nbits), q_codes.T.astype(np.uint8))
            assert np.array_equal(O.gptq_unpack_fast(layer.qweight.numpy(), nbits), q_codes.T.astype(np.uint8))

            # forward (torch dequant+matmul branch), M in {1, 5}
            fx = {}
            for M in (1, 5):
                torch.manual_seed(7 + M)
                x = torch.randn(M, K).half()
                y_ref = layer(x)
                y_or = O.gptq_forward_torch(x, layer.qweight.numpy(), layer.scales, layer.zeros, nbits, G)
                assert torch.equal(y_ref, y_or), ("gptq fwd", nbits, M)
                y32 = O.gptq_forward_fp32(x, layer.qweight.numpy(), layer.scales, layer.zeros, nbits, G)
                fx[f"x{M}"] = x.numpy()
                fx[f"y{M}_ref_fp16"] = y_ref.numpy()
                fx[f"y{M}_fp32"] = y32.numpy()
                report.append(f"N{N} K{K} b{nbits} M{M}: ref-fp16 vs fp32 max-rel {O.max_rel(y_ref, y32):.2e}")

            extra = {}
            if nbits == 4:
                ft_q = pack_intweight(torch.from_numpy(q_codes.astype(np.int32)), interleave=4, kstride=64).numpy()
                assert np.array_equal(O.ft_pack_intweight(q_codes), ft_q), "ft pack"
                assert np.array_equal(O.ft_unpack(ft_q), q_codes.astype(np.uint8)), "ft unpack"
                extra["ft_qweight"] = ft_q
            np.savez_compressed(
                os.path.join(out_dir, f"linear_{nbits}bit_N{N}_K{K}.npz"),
                W=W.numpy(), hqq_Wq=W_q.numpy(), hqq_scale=meta["scale"].numpy(),
                hqq_zero=meta["zero"].numpy(), codes=codes, W_deq=W_deq.numpy(),
                gptq_qweight=layer.qweight.numpy(), gptq_scales=layer.scales.numpy(),
                gptq_zeros=layer.zeros.numpy(), solver_iters=np.int32(n_it), **fx, **extra)
    report.append("quantize/dequant/gptq/ft ok")

    hqqlinear_state_fixtures(out_dir, report)
    fp16_solver_fixtures(out_dir, report)
    proxy_checkpoint_fixture(out_dir, report)
    hf_dropin_fixture(out_dir, report)

    # ---- 3. arch selection rule (amq_speed_benchmark.py:209-229) on a synthetic stats file
    import json
    rs = np.random.RandomState(0)
    names = ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
             "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"]
    entries = []
    for i in range(12):
        arch = {"linear": {n: rs.choice([2, 3, 4], size=4).tolist() for n in names}}
        entries.append([arch, float(rs.rand()), float(2.9 + 0.02 * i)])
    stats = {"archive": entries[:8], "candidates": entries[8:]}
    # reference rule, executed verbatim in spirit: filter, count 4-bit, argmax
    cands = [a for a in stats["archive"] + stats["candidates"] if abs(a[-1] - 3.0) < 0.05]
    cb = [np.concatenate([b for b in a[0]["linear"].values()]) for a in cands]
    pick = cands[int(np.argmax([(b == 4.0).sum() for b in cb]))][0]["linear"]
    assert O.select_arch(stats, 3.0) == pick
    with open(os.path.join(out_dir, "arch_stats.json"), "w") as f:
        json.dump({"stats": stats, "target_bits": 3.0, "expected": pick}, f)
    report.append("arch selection ok")

    print("\n".join(report))


if __name__ == "__main__":
    main()
