"""CPU oracle for the AMQ quantized-linear hot path.

TEST INFRASTRUCTURE ONLY.  This file is a CPU restatement (numpy for the
integer / byte work, torch-CPU for the fp16/fp32 arithmetic whose rounding
must match the reference's torch ops) of the reference algorithms on the hot
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; nothing under
``amq_b200/`` does.  The product path has no CPU fallback.

Parity pin: ``oracle/gen_golden.py`` imports the *unmodified* reference
(``/root/reference/amq/kernel/hqq``) in the build container, checks every
function below against it on seeded inputs and writes the fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks the oracle
against those fixtures everywhere (the reference itself does not travel).
The reference ships no golden vectors of its own for this path.

Reference citations are relative to /root/reference/ with
HQQ = amq/kernel/hqq/hqq.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------
# HQQ bit packing (HQQ/core/bitpack.py:24-110): block-strided over rows.
# --------------------------------------------------------------------------
_HQQ_FIELDS = {4: 2, 2: 4, 3: 10, 1: 8, 8: 1}  # codes per packed element


def hqq_pack(codes: np.ndarray, nbits: int) -> np.ndarray:
    """codes [R, G] (values < 2**nbits) -> packed W_q.

    bitpack.py:24-28 (4-bit u8), :43-52 (2-bit u8), :69-91 (3-bit int32, rows
    zero-padded to a multiple of 10, field j at shift 3*(9-j))."""
    codes = np.asarray(codes)
    R, G = codes.shape
    p = _HQQ_FIELDS[nbits]
    if nbits == 8:
        return codes.astype(np.uint8)
    if nbits == 3:
        Rp = 10 * int(math.ceil(R / 10.0))
        buf = np.zeros((Rp, G), dtype=np.int32)
        buf[:R] = codes.astype(np.int32)
        step = Rp // 10
        out = np.zeros((step, G), dtype=np.int32)
        for j in range(10):
            out |= buf[j * step:(j + 1) * step] << (3 * (9 - j))
        return out
    # u8 containers; the reference uses int(len/p) (truncating) block size
    step = int(R / p)
    c = codes.astype(np.uint8)
    out = np.zeros((step, G), dtype=np.uint8)
    for j in range(p):
        blk = c[j * step:(j + 1) * step] if j < p - 1 else c[(p - 1) * step:]
        out |= (blk[:step] << (nbits * (p - 1 - j))).astype(np.uint8)
    return out


def hqq_unpack(W_q: np.ndarray, nbits: int, rows: Optional[int] = None) -> np.ndarray:
    """Inverse of hqq_pack (bitpack.py:30-38, 54-64, 95-110).  ``rows`` trims
    the 3-bit padding like quantize.py:190-195 does."""
    W_q = np.asarray(W_q)
    if nbits == 8:
        return W_q.astype(np.uint8)
    p = _HQQ_FIELDS[nbits]
    step, G = W_q.shape
    out = np.empty((p * step, G), dtype=np.uint8)
    mask = (1 << nbits) - 1
    wide = W_q.astype(np.int64) & 0xFFFFFFFF
    for j in range(p):
        out[j * step:(j + 1) * step] = ((wide >> (nbits * (p - 1 - j))) & mask).astype(np.uint8)
    if rows is not None:
        out = out[:rows]
    return out


# --------------------------------------------------------------------------
# HQQ quantizer (HQQ/core/quantize.py:75-180, HQQ/core/optimize.py:96-108,
# 201-255).  axis=1 only (AMQ's proxies, amq_quantization_proxy.py:36).
# --------------------------------------------------------------------------
def _shrink_lp(x: torch.Tensor, beta: float, p: float) -> torch.Tensor:
    # optimize.py:96-108
    a = x.abs()
    if p == 1:
        mag = (a - 1.0 / beta).clamp_min(0.0)
    else:
        mag = (a - (1.0 / beta) * a.pow(p - 1)).clamp_min(0.0)
    return mag * torch.sign(x)


def hqq_quantize(W: torch.Tensor, nbits: int, group_size: int = 128,
                 round_zero: Optional[bool] = None, optimize: bool = True,
                 solver_dtype: torch.dtype = torch.float32,
                 iters: int = 20, beta: float = 10.0, lp_norm: float = 0.7,
                 ) -> Tuple[np.ndarray, torch.Tensor, torch.Tensor, int]:
    """Returns (codes [R,G] uint8, scale [R,1], zero [R,1], iterations_run).

    scale is already inverted (dequant scale), as stored in meta['scale']
    (quantize.py:154).  ``round_zero`` defaults to nbits == 4
    (quantize.py:1097).  ``solver_dtype`` fp32 = the reference's CPU branch,
    fp16 = its CUDA branch (optimize.py:231)."""
    if round_zero is None:
        round_zero = nbits == 4
    Wf = W.detach().to("cpu").float().reshape(-1, group_size)
    mn = Wf.min(dim=1, keepdim=True)[0]
    mx = Wf.max(dim=1, keepdim=True)[0]
    max_v = float(round(2 ** nbits - 1))
    denom = mx - mn
    scale = max_v / denom
    scale = torch.where(denom.abs() <= 1e-4, torch.full_like(scale, 1.0), scale)
    scale = scale.clamp(max=2e4)
    zero = -mn * scale
    if round_zero:
        zero = torch.round(zero)
    n_it = 0
    if optimize:
        Ws = Wf.to(solver_dtype)
        s = scale.to(solver_dtype)
        z = zero.to(solver_dtype)
        best = torch.tensor(float("inf"), dtype=torch.float32)
        for _ in range(iters):
            n_it += 1
            Wq = torch.round(Ws * s + z).clamp_(0.0, max_v)
            Wr = (Wq - z) / s
            We = _shrink_lp(Ws - Wr, beta, lp_norm)
            z = torch.mean(Wq - (Ws - We) * s, dim=1, keepdim=True)
            err = torch.abs(Ws - Wr).mean().float()
            if err < best:
                best = err
            else:
                break
        scale, zero = s, z
        # final codes from the ORIGINAL fp32 tensor (optimize.py:254); torch
        # type promotion makes the product fp32 even when s/z are fp16.
        Wq = torch.round(Wf * scale + zero).clamp_(0.0, max_v)
    else:
        Wq = (Wf * scale + zero).round_().clamp_(0.0, max_v)
    inv = 1.0 / scale
    return Wq.to(torch.uint8).numpy(), inv, zero, n_it


def hqq_dequantize(codes: np.ndarray, scale: torch.Tensor, zero: torch.Tensor,
                   shape: Sequence[int], dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """quantize.py:183-199: ((W_r - zero) * scale).reshape(shape) computed in
    ``dtype`` (two roundings in fp16)."""
    Wr = torch.from_numpy(np.asarray(codes)).to(dtype)
    return ((Wr - zero.to(dtype)) * scale.to(dtype)).reshape(tuple(shape))


# --------------------------------------------------------------------------
# GPTQLinear layout (HQQ/backends/autogptq.py:55-82, 111-156, 245-283).
# --------------------------------------------------------------------------
def gptq_codes_from_weight(W: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor,
                           group_size: int) -> np.ndarray:
    """autogptq.py:112-121: q = round((W + zero*scale)/scale), in the dtype of
    the inputs (fp16 in the AMQ flow).  Returns codes [N, K] int64."""
    sz = zeros * scales
    s_i = torch.repeat_interleave(scales, group_size, dim=1)
    sz_i = torch.repeat_interleave(sz, group_size, dim=1)
    q = torch.round((W + sz_i) / s_i).to(torch.int)
    return q.numpy().astype(np.int64)


def gptq_pack_codes(codes: np.ndarray, bits: int) -> np.ndarray:
    """codes [N, K] -> qweight int32 [K*bits/32, N].

    For every output column the K codes form one little-endian bit stream of
    ``bits`` bits per code, cut into 32-bit words.  This is
    what the hand-written case analysis at autogptq.py:124-153 produces
    (3-bit: code 10 split 2+1, code 21 split 1+2)."""
    codes = np.asarray(codes).astype(np.uint64)
    N, K = codes.shape
    assert (K * bits) % 32 == 0
    rows = K * bits // 32
    q = np.zeros((rows, N), dtype=np.uint64)
    ct = codes.T  # [K, N]
    for k in range(K):
        pos = k * bits
        r, sh = divmod(pos, 32)
        q[r] |= (ct[k] << np.uint64(sh)) & np.uint64(0xFFFFFFFF)
        if sh + bits > 32:
            q[r + 1] |= ct[k] >> np.uint64(32 - sh)
    return q.astype(np.uint32).view(np.int32)


def gptq_unpack(qweight: np.ndarray, bits: int) -> np.ndarray:
    """qweight int32 [K*bits/32, N] -> codes uint8 [K, N]
    (autogptq.py:256-261 for 2/4/8-bit, :269-277 for 3-bit)."""
    q = np.asarray(qweight).view(np.uint32).astype(np.uint64)
    rows, N = q.shape
    K = rows * 32 // bits
    out = np.empty((K, N), dtype=np.uint8)
    mask = np.uint64((1 << bits) - 1)
    for k in range(K):
        pos = k * bits
        r, sh = divmod(pos, 32)
        v = q[r] >> np.uint64(sh)
        if sh + bits > 32:
            v = v | (q[r + 1] << np.uint64(32 - sh))
        out[k] = (v & mask).astype(np.uint8)
    return out


def gptq_unpack_fast(qweight: np.ndarray, bits: int) -> np.ndarray:
    """Vectorised gptq_unpack (same result; used for model-scale tensors)."""
    q = np.ascontiguousarray(qweight).view(np.uint32)
    rows, N = q.shape
    if bits in (2, 4, 8):
        per = 32 // bits
        sh = (np.arange(per, dtype=np.uint32) * bits)[None, :, None]
        out = (q[:, None, :] >> sh) & np.uint32((1 << bits) - 1)
        return out.reshape(rows * per, N).astype(np.uint8)
    assert bits == 3
    q3 = q.reshape(rows // 3, 3, N).astype(np.uint64)
    # 96-bit streams do not fit uint64: split at code 21 (stream bits 63..65)
    lo = q3[:, 0] | (q3[:, 1] << np.uint64(32))          # stream bits 0..63
    hi = (q3[:, 1] >> np.uint64(31)) | (q3[:, 2] << np.uint64(1))  # stream bits 63..95 (+)
    out = np.empty((rows // 3, 32, N), dtype=np.uint8)
    for j in range(32):
        pos = 3 * j
        if pos + 3 <= 64:
            out[:, j] = ((lo >> np.uint64(pos)) & np.uint64(7)).astype(np.uint8)
        else:
            out[:, j] = ((hi >> np.uint64(pos - 63)) & np.uint64(7)).astype(np.uint8)
    return out.reshape(rows // 3 * 32, N)


def gptq_dequant_fp16(qweight: np.ndarray, scales: torch.Tensor, zeros: torch.Tensor,
                      bits: int, group_size: int) -> torch.Tensor:
    """autogptq.py:246-282: weight[K,N] = scales.half()*q - zeros.half(), fp16
    (int8 codes promoted to fp16 by torch; two fp16 roundings)."""
    codes = torch.from_numpy(gptq_unpack_fast(qweight, bits).astype(np.int8))
    K, N = codes.shape
    s = scales.half().reshape(-1, 1, N)
    z = zeros.half().reshape(-1, 1, N)
    w = s * codes.reshape(-1, group_size, N) - z
    return w.reshape(K, N)


def gptq_forward_torch(x: torch.Tensor, qweight: np.ndarray, scales: torch.Tensor,
                       zeros: torch.Tensor, bits: int, group_size: int,
                       bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The reference's torch dequant+matmul branch, GPTQLinear.forward with
    kernel_switch_threshold=0 (autogptq.py:159-163, 245-288), on CPU."""
    out_shape = x.shape[:-1] + (scales.shape[1],)
    x2 = x.reshape(-1, x.shape[-1])
    w = gptq_dequant_fp16(qweight, scales, zeros, bits, group_size)
    out = torch.matmul(x2, w.to(x2.dtype)).to(x.dtype).reshape(out_shape)
    return out + bias if bias is not None else out


def gptq_forward_fp32(x: torch.Tensor, qweight: np.ndarray, scales: torch.Tensor,
                      zeros: torch.Tensor, bits: int, group_size: int,
                      bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Exact-fp32 reference for the <=1e-3 tolerance:
    x.float() @ (scales.float()*q - zeros.float()), from the STORED buffers."""
    codes = torch.from_numpy(gptq_unpack_fast(qweight, bits).astype(np.float32))
    K, N = codes.shape
    w = (scales.float().reshape(-1, 1, N) * codes.reshape(-1, group_size, N)
         - zeros.float().reshape(-1, 1, N)).reshape(K, N)
    out = x.reshape(-1, K).float() @ w
    if bias is not None:
        out = out + bias.float()
    return out.reshape(x.shape[:-1] + (N,))


# --------------------------------------------------------------------------
# FT_QuantLinear layout (HQQ/backends/ft.py:15-55, 103-126).
# --------------------------------------------------------------------------
def ft_pack_intweight(codes: np.ndarray, interleave: int = 4, kstride: int = 64) -> np.ndarray:
    """codes [N, K] (4-bit) -> int16 [N/4, K] (ft.py:15-55).  Written as the
    closed-form index map of SURVEY App. A3 rather than the reshape chain."""
    codes = np.asarray(codes).astype(np.uint16)
    N, K = codes.shape
    assert interleave == 4 and kstride == 64 and N % 4 == 0 and K % 64 == 0
    n = np.arange(N)[:, None]
    k = np.arange(K)[None, :]
    off = k % 32
    # off = 8*(pos % 4) + pos // 4 after the first reorder, then each 8-run
    # [0..7] -> [0,2,4,6,1,3,5,7].
    # position after reorder 1: inverse of pos -> off
    #   pos1 = 4*(off % 8) + off // 8      is wrong; derive by table instead.
    tbl1 = np.arange(32).reshape(4, 4, 2).transpose(1, 0, 2).reshape(32)   # pos1 -> off
    tbl2 = np.arange(32).reshape(4, 4, 2).transpose(0, 2, 1).reshape(32)   # pos2 -> pos1
    src = tbl1[tbl2]                      # pos2 -> original offset inside the 32-chunk
    inv = np.empty(32, dtype=np.int64)
    inv[src] = np.arange(32)              # original offset -> pos2
    pos = inv[off]
    kk = (32 * (k // 32) + pos) % 64
    tile = k // 64
    t = (n % 4) * 64 + kk                 # index inside the [4 rows x 64] tile
    col = 64 * tile + t // 4
    nib = t % 4
    out = np.zeros((N // 4, K), dtype=np.uint16)
    np.bitwise_or.at(out, (np.broadcast_to(n // 4, (N, K)), col), (codes << (4 * nib)).astype(np.uint16))
    return out.view(np.int16)


def ft_unpack(qweight: np.ndarray) -> np.ndarray:
    """int16 [N/4, K] -> codes uint8 [N, K]; inverse of ft_pack_intweight (the
    order dequantize_s4_to_fp16x2 + gemv_kernel consume,
    ft/quantization_new/dequantize.cuh:14-77, gemv/gemv_cuda.cu:73-204)."""
    q = np.asarray(qweight).view(np.uint16)
    N4, K = q.shape
    N = N4 * 4
    n = np.arange(N)[:, None]
    k = np.arange(K)[None, :]
    tbl1 = np.arange(32).reshape(4, 4, 2).transpose(1, 0, 2).reshape(32)
    tbl2 = np.arange(32).reshape(4, 4, 2).transpose(0, 2, 1).reshape(32)
    src = tbl1[tbl2]
    inv = np.empty(32, dtype=np.int64)
    inv[src] = np.arange(32)
    pos = inv[k % 32]
    kk = (32 * (k // 32) + pos) % 64
    t = (n % 4) * 64 + kk
    col = 64 * (k // 64) + t // 4
    nib = t % 4
    return ((q[np.broadcast_to(n // 4, (N, K)), col] >> (4 * nib)) & 0xF).astype(np.uint8)


def ft_forward_fp32(x: torch.Tensor, qweight: np.ndarray, scales: torch.Tensor,
                    scaled_zeros: torch.Tensor, group_size: int,
                    bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """W[n,k] = q*scales[g,n] + scaled_zeros[g,n] (ft.py:121-126: scaled_zeros
    = -(zero*scale)), exact fp32 from the stored buffers."""
    codes = torch.from_numpy(ft_unpack(qweight).astype(np.float32))   # [N, K]
    N, K = codes.shape
    s = scales.float().t().reshape(N, -1, 1)
    z = scaled_zeros.float().t().reshape(N, -1, 1)
    w = (codes.reshape(N, -1, group_size) * s + z).reshape(N, K)
    out = x.reshape(-1, K).float() @ w.t()
    if bias is not None:
        out = out + bias.float()
    return out.reshape(x.shape[:-1] + (N,))


# --------------------------------------------------------------------------
# Search-output bit config (amq/amq_speed_benchmark.py:209-229,
# amq/utils/func.py:101-114).
# --------------------------------------------------------------------------
def get_bits_usage(arch: Dict, config: Dict, group_size: int = 128) -> float:
    mem = 0.0
    for name, bits in arch["linear"].items():
        out_dim, in_dim = config["linear_shape"][name]
        g = in_dim if group_size == -1 else group_size
        for b in bits:
            eff = b + (32 / g if b < 16 else 0)
            mem += int(out_dim) * int(in_dim) * eff
    return mem / config["model_numel"]


def select_arch(stats: Dict, target_bits: float) -> Dict[str, List[int]]:
    """amq_speed_benchmark.py:209-229: among archive+candidates entries whose
    bits_usage is within 0.05 of the target take the one with the most 4-bit
    linears (first on ties, np.argmax)."""
    archs = stats["archive"] + stats["candidates"]
    cands = [a for a in archs if abs(a[-1] - target_bits) < 0.05]
    if not cands:
        raise ValueError("no candidate within 0.05 bits of target")
    counts = [int(sum(int(b == 4.0) for bits in a[0]["linear"].values() for b in bits)) for a in cands]
    return cands[int(np.argmax(counts))][0]["linear"]


def max_rel(y: torch.Tensor, ref: torch.Tensor) -> float:
    """SURVEY §8d parity metric: max|y - ref| / max|ref|."""
    return float((y.float() - ref.float()).abs().max() / ref.float().abs().max())
