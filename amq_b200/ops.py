"""Python shims over the C ABI (include/amqb.h).  The shims own all allocations (torch tensors)
and pass raw device pointers + the current CUDA stream; the library allocates nothing."""
from __future__ import annotations

import ctypes
import functools
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import GemvProblem, check, cur_stream, lib, ptr

GROUP = 128
_workspaces: Dict[Tuple[int, int], torch.Tensor] = {}


def _on_device(fn):
    """Run the op with the device of its first CUDA tensor argument current: the launch stream, the workspace and
    the library's per-device state all follow the TENSORS, not whatever device happens to be current (the reference
    guards the same way, OptionalCUDAGuard(device_of(vec)), auto_gptq_kernel.cu:448)."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for v in list(args) + list(kwargs.values()):
            if isinstance(v, torch.Tensor) and v.is_cuda:
                if v.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(v.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapper


def native_supported(bits: int, N: int, K: int, G: int) -> bool:
    return bits in (2, 3, 4) and G == GROUP and N % 32 == 0 and K % GROUP == 0 and N > 0 and K > 0


def native_bytes(bits: int, N: int, K: int) -> int:
    return int(lib().amqb_native_bytes(bits, N, K))


def workspace(device: torch.device, sum_N: int = 32768, max_K: int = 16384, max_M: int = 1) -> torch.Tensor:
    """Zero-filled split-K workspace, one per (device, stream), grown on demand (never while a CUDA
    graph is being captured: size it first with the largest launch).  The decode kernels leave the
    arrival counters zeroed, so it is cleared exactly once per allocation."""
    dev = device.index if device.index is not None else torch.cuda.current_device()
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ws = _workspaces.get(key)
    need = max(int(lib().amqb_workspace_bytes(sum_N, max_K, max_M)), 256)
    if need == 0:
        raise RuntimeError("amq_b200: workspace request out of range")
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=torch.device("cuda", dev))
        _workspaces[key] = ws
    return ws


def _req_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("amq_b200: CUDA tensors required (no CPU fallback on this path)")


# ------------------------------------------------------------------ layout transcoders
@_on_device
def unpack_codes(packed: torch.Tensor, bits: int, layout: int, N: int, K: int, G: int = GROUP) -> torch.Tensor:
    _req_cuda(packed)
    out = torch.empty((N, K), dtype=torch.uint8, device=packed.device)
    check(lib().amqb_unpack_codes(bits, layout, ptr(packed), ptr(out), N, K, G, cur_stream()), "unpack_codes")
    return out


@_on_device
def repack_gptq(bits: int, qweight: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor,
                N: int, K: int, G: int = GROUP) -> torch.Tensor:
    _req_cuda(qweight, scales, zeros)
    nat = torch.empty(native_bytes(bits, N, K), dtype=torch.uint8, device=qweight.device)
    scratch = torch.empty((N, K), dtype=torch.uint8, device=qweight.device)
    # converted copies are bound to locals: a temporary inside the call expression is freed as soon as ptr() returns
    # and the caching allocator hands the same block to the next same-sized temporary
    qw, sc, zs = qweight.contiguous(), scales.float().contiguous(), zeros.float().contiguous()
    check(lib().amqb_repack_gptq(bits, ptr(qw), ptr(sc), ptr(zs), ptr(nat), ptr(scratch), N, K, G, cur_stream()), "repack_gptq")
    return nat


@_on_device
def repack_ft(qweight: torch.Tensor, scales: torch.Tensor, scaled_zeros: torch.Tensor,
              N: int, K: int, G: int = GROUP) -> torch.Tensor:
    _req_cuda(qweight, scales, scaled_zeros)
    nat = torch.empty(native_bytes(4, N, K), dtype=torch.uint8, device=qweight.device)
    scratch = torch.empty((N, K), dtype=torch.uint8, device=qweight.device)
    qw, sc, zs = qweight.contiguous(), scales.half().contiguous(), scaled_zeros.half().contiguous()
    check(lib().amqb_repack_ft(ptr(qw), ptr(sc), ptr(zs), ptr(nat), ptr(scratch), N, K, G, cur_stream()), "repack_ft")
    return nat


@_on_device
def pack_native(bits: int, codes: torch.Tensor, scale: torch.Tensor, zero: torch.Tensor,
                zero_is_scaled: bool = False, G: int = GROUP) -> torch.Tensor:
    """codes u8 [N,K]; scale/zero fp16 [N, K/G] (HQQ meta; W = (q - zero) * scale)."""
    _req_cuda(codes, scale, zero)
    N, K = codes.shape
    nat = torch.empty(native_bytes(bits, N, K), dtype=torch.uint8, device=codes.device)
    cd, sc, zr = codes.contiguous(), scale.half().contiguous(), zero.half().contiguous()
    check(lib().amqb_pack_native(bits, ptr(cd), ptr(sc), ptr(zr), int(zero_is_scaled), ptr(nat), N, K, G, cur_stream()),
          "pack_native")
    return nat


@_on_device
def gptq_pack(bits: int, W: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor, G: int):
    """GPTQLinear.pack on the GPU. W fp16 [N,K]; scales/zeros [N,K/G] -> (qweight, scales_f32, zeros_f32)."""
    _req_cuda(W, scales, zeros)
    N, K = W.shape
    qweight = torch.empty((K // 32 * bits, N), dtype=torch.int32, device=W.device)
    s_out = torch.empty((K // G, N), dtype=torch.float32, device=W.device)
    z_out = torch.empty((K // G, N), dtype=torch.float32, device=W.device)
    Wh, sc, zr = W.half().contiguous(), scales.half().contiguous(), zeros.half().contiguous()
    check(lib().amqb_gptq_pack(bits, ptr(Wh), ptr(sc), ptr(zr), ptr(qweight), ptr(s_out), ptr(z_out), N, K, G,
                               cur_stream()), "gptq_pack")
    return qweight, s_out, z_out


@_on_device
def ft_pack(W: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor, G: int):
    _req_cuda(W, scales, zeros)
    N, K = W.shape
    qweight = torch.empty((N // 4, K), dtype=torch.int16, device=W.device)
    s_out = torch.empty((K // G, N), dtype=torch.float16, device=W.device)
    z_out = torch.empty((K // G, N), dtype=torch.float16, device=W.device)
    Wh, sc, zr = W.half().contiguous(), scales.half().contiguous(), zeros.half().contiguous()
    check(lib().amqb_ft_pack(ptr(Wh), ptr(sc), ptr(zr), ptr(qweight), ptr(s_out), ptr(z_out), N, K, G, cur_stream()), "ft_pack")
    return qweight, s_out, z_out


# ------------------------------------------------------------------ decode
def make_problem(bits: int, w_native: torch.Tensor, x: torch.Tensor, y: torch.Tensor, N: int, K: int,
                 bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
                 prologue: int = _lib.PRO_NONE, gamma: Optional[torch.Tensor] = None, eps: float = 0.0,
                 ldx: Optional[int] = None, ldy: Optional[int] = None) -> GemvProblem:
    M = x.shape[0]
    p = GemvProblem()
    p.bits, p.M, p.N, p.K = bits, M, N, K
    p.w_native = w_native.data_ptr()
    p.x = x.data_ptr()
    p.ldx = ldx if ldx is not None else x.stride(0)
    p.y = y.data_ptr()
    p.ldy = ldy if ldy is not None else y.stride(0)
    p.bias = bias.data_ptr() if bias is not None else None
    p.residual = residual.data_ptr() if residual is not None else None
    p.prologue = prologue
    p.gamma = gamma.data_ptr() if gamma is not None else None
    p.eps = eps
    return p


def gemv_grouped(problems: Sequence[GemvProblem], ws: torch.Tensor, pdl: bool = False) -> None:
    arr = (GemvProblem * len(problems))(*problems)
    check(lib().amqb_gemv_grouped(arr, len(problems), ptr(ws), ctypes.c_size_t(ws.numel()), int(pdl), cur_stream()),
          "gemv_grouped")


@_on_device
def gemv(bits: int, w_native: torch.Tensor, x: torch.Tensor, N: int, K: int,
         bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y fp16 [M,N] = x[M,K] @ dequant(W)^T (+bias), M <= 16, through amqb_gemv_w{2,3,4}."""
    _req_cuda(w_native, x)
    if x.dtype != torch.float16:
        raise RuntimeError("amq_b200: fp16 activations required (as the reference kernels, autogptq.py:165-169, ft.py:62)")
    x = x.contiguous()
    M = x.shape[0]
    y = out if out is not None else torch.empty((M, N), dtype=torch.float16, device=x.device)
    ws = workspace(x.device, N, K, M)
    fn = {2: lib().amqb_gemv_w2, 3: lib().amqb_gemv_w3, 4: lib().amqb_gemv_w4}[bits]
    check(fn(ptr(w_native), ptr(x), ptr(y), ptr(bias), M, N, K, ptr(ws), ctypes.c_size_t(ws.numel()), cur_stream()),
          f"gemv_w{bits}")
    return y


@_on_device
def gemv_gptq_layout(bits: int, qweight: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor, x: torch.Tensor,
                     N: int, K: int, G: int, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req_cuda(qweight, scales, zeros, x)
    x = x.contiguous()
    M = x.shape[0]
    y = torch.empty((M, N), dtype=torch.float16, device=x.device)
    check(lib().amqb_gemv_gptq_layout(bits, ptr(qweight), ptr(scales), ptr(zeros), ptr(x), ptr(y), ptr(bias),
                                      M, N, K, G, cur_stream()), "gemv_gptq_layout")
    return y


def linear_forward(bits: int, w_native: torch.Tensor, x2: torch.Tensor, N: int, K: int,
                   bias: Optional[torch.Tensor]) -> torch.Tensor:
    """Dispatch on the row count like the reference modules do (autogptq.py:163, ft.py:129-142):
    decode kernel for M <= 16, tensor-core prefill kernel above."""
    M = x2.shape[0]
    if M <= 16:
        return gemv(bits, w_native, x2, N, K, bias)
    return gemm_tc(bits, w_native, x2, N, K, bias)


def gemm_workspace(M: int, K: int, bits: int, device) -> torch.Tensor:
    """Scratch for `gemm_tc` (the pre-swizzled activations); reusable across calls of the same or smaller M, K."""
    need = int(lib().amqb_gemm_workspace_bytes(M, K, bits))
    return torch.empty(max(need, 256), dtype=torch.uint8, device=device)


@_on_device
def gemm_tc(bits: int, w_native: torch.Tensor, x: torch.Tensor, N: int, K: int,
            bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
            workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req_cuda(w_native, x)
    x = x.contiguous()
    M = x.shape[0]
    y = out if out is not None else torch.empty((M, N), dtype=torch.float16, device=x.device)
    if y.shape != (M, N) or y.dtype != torch.float16 or not y.is_contiguous():
        raise ValueError("gemm_tc: out must be a contiguous fp16 [M, N] tensor")
    L = lib()
    if not hasattr(L, "amqb_gemm_tc"):
        raise RuntimeError("amq_b200: amqb_gemm_tc missing from libamqb.so")
    wsb = workspace if workspace is not None else gemm_workspace(M, K, bits, x.device)
    check(L.amqb_gemm_tc(bits, ptr(w_native), ptr(x), ptr(y), ptr(bias), M, N, K, ptr(wsb),
                         ctypes.c_size_t(wsb.numel()), cur_stream()), "gemm_tc")
    return y


@_on_device
def linear_forward_grouped(members, x: torch.Tensor, K: int, ws: Optional[torch.Tensor] = None):
    """Several linears over the same activations x [M, K] (q|k|v, gate|up): members = [(bits, w_native, N, bias or None)].
    More than 16 rows: ONE tcgen05 launch (amqb_gemm_tc_grouped); up to 16: the decode kernel's grouped launch.  Returns
    the list of outputs [M, N_i]."""
    from ._lib import GemmProblem
    x = x.contiguous()
    M = x.shape[0]
    outs = [torch.empty((M, N), dtype=torch.float16, device=x.device) for (_, _, N, _) in members]
    if M <= 16:
        wsd = ws if ws is not None else workspace(x.device, sum(m[2] for m in members), K, M)
        probs = [make_problem(b, w, x, y, N, K, bias=bias) for (b, w, N, bias), y in zip(members, outs)]
        gemv_grouped(probs, wsd)
        return outs
    bits_max = max(m[0] for m in members)
    wsb = ws if ws is not None else gemm_workspace(M, K, bits_max, x.device)
    arr = (GemmProblem * len(members))()
    keep = []
    for i, ((b, w, N, bias), y) in enumerate(zip(members, outs)):
        _req_cuda(w)
        arr[i].bits, arr[i].N = b, N
        arr[i].w_native, arr[i].y = w.data_ptr(), y.data_ptr()
        arr[i].bias = bias.data_ptr() if bias is not None else None
        keep.append((w, bias))
    check(lib().amqb_gemm_tc_grouped(arr, len(members), ptr(x), M, K, ptr(wsb), ctypes.c_size_t(wsb.numel()), cur_stream()),
          "gemm_tc_grouped")
    return outs


# ------------------------------------------------------------------ HQQ proxy ops
@_on_device
def hqq_dequant(bits: int, W_q: torch.Tensor, scale: torch.Tensor, zero: torch.Tensor, N: int, K: int,
                G: int = GROUP) -> torch.Tensor:
    _req_cuda(W_q, scale, zero)
    out = torch.empty((N, K), dtype=torch.float16, device=W_q.device)
    wq, sc, zr = W_q.contiguous(), scale.half().contiguous(), zero.half().contiguous()
    check(lib().amqb_hqq_dequant(bits, ptr(wq), ptr(sc), ptr(zr), ptr(out), N, K, G, cur_stream()), "hqq_dequant")
    return out


@_on_device
def hqq_pack(bits: int, codes: torch.Tensor) -> torch.Tensor:
    _req_cuda(codes)
    R, G = codes.shape
    if bits == 3:
        out = torch.empty(((R + 9) // 10, G), dtype=torch.int32, device=codes.device)
    else:
        out = torch.empty((R // (2 if bits == 4 else 4), G), dtype=torch.uint8, device=codes.device)
    cd = codes.to(torch.uint8).contiguous()
    check(lib().amqb_hqq_pack(bits, ptr(cd), ptr(out), R, G, cur_stream()), "hqq_pack")
    return out


@_on_device
def hqq_unpack(bits: int, W_q: torch.Tensor, R: Optional[int] = None) -> torch.Tensor:
    _req_cuda(W_q)
    step, G = W_q.shape
    p = {4: 2, 2: 4, 3: 10}[bits]
    Rfull = step * p
    out = torch.empty((Rfull, G), dtype=torch.uint8, device=W_q.device)
    wq = W_q.contiguous()
    check(lib().amqb_hqq_unpack(bits, ptr(wq), ptr(out), Rfull, G, cur_stream()), "hqq_unpack")
    return out if R is None else out[:R]


@_on_device
def hqq_quantize(W: torch.Tensor, bits: int, G: int = GROUP, round_zero: Optional[bool] = None,
                 solver_dtype: torch.dtype = torch.float32, packed: bool = False, want_codes: bool = True):
    """Quantizer.quantize (axis=1) -> (codes u8 [R,G] or None, scale fp32 [R,1], zero fp32 [R,1], iters[, W_q]).
    solver_dtype: torch.float32 = the arithmetic of the reference's CPU branch (bit-exact with it), torch.float16 = its
    CUDA branch (optimize.py:231: every solver op rounded to fp16).  packed: also return HQQ's packed W_q (BitPack
    layout), written in the same pass as the codes."""
    _req_cuda(W)
    if round_zero is None:
        round_zero = bits == 4
    N, K = W.shape
    R = N * K // G
    L = lib()
    codes = torch.empty((R, G), dtype=torch.uint8, device=W.device) if (want_codes or not packed) else None
    scale = torch.empty((R, 1), dtype=torch.float32, device=W.device)
    zero = torch.empty((R, 1), dtype=torch.float32, device=W.device)
    iters = torch.zeros(1, dtype=torch.int32, device=W.device)
    need = int(L.amqb_hqq_quantize_workspace_bytes(N, K, G))
    wsb = torch.empty(max(need, 16), dtype=torch.uint8, device=W.device)
    Wh = W.half().contiguous()
    if not packed:
        if solver_dtype != torch.float32:
            raise ValueError("hqq_quantize: the fp16 solver is served by the packed entry point (packed=True)")
        check(L.amqb_hqq_quantize(bits, ptr(Wh), ptr(codes), ptr(scale), ptr(zero), int(round_zero),
                                  N, K, G, ptr(wsb), ctypes.c_size_t(wsb.numel()), ptr(iters), cur_stream()), "hqq_quantize")
        return codes, scale, zero, iters
    if bits == 3:
        W_q = torch.empty(((R + 9) // 10, G), dtype=torch.int32, device=W.device)
    else:
        W_q = torch.empty((R // (2 if bits == 4 else 4), G), dtype=torch.uint8, device=W.device)
    check(L.amqb_hqq_quantize_packed(bits, ptr(Wh), ptr(W_q), ptr(codes), ptr(scale), ptr(zero), int(round_zero),
                                     int(solver_dtype == torch.float16), N, K, G, ptr(wsb), ctypes.c_size_t(wsb.numel()),
                                     ptr(iters), cur_stream()), "hqq_quantize_packed")
    return codes, scale, zero, iters, W_q


@_on_device
def native_to_dense(bits: int, w_native: torch.Tensor, N: int, K: int) -> torch.Tensor:
    """fp32 [N, K] weights a native buffer encodes: scale * q - zero*scale (test / debugging aid)."""
    codes = unpack_codes(w_native, bits, _lib.LAYOUT_NATIVE, N, K, GROUP).float()
    rec = bits * 512 + 128
    meta = w_native.reshape(-1, rec)[:, bits * 512:].contiguous().view(torch.float16).reshape(N // 32, K // GROUP, 32, 2)
    scale = meta[..., 0].permute(0, 2, 1).reshape(N, K // GROUP).float()
    zs = meta[..., 1].permute(0, 2, 1).reshape(N, K // GROUP).float()
    return (codes.reshape(N, K // GROUP, GROUP) * scale[..., None] - zs[..., None]).reshape(N, K)
