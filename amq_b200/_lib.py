"""ctypes binding of libamqb.so (include/amqb.h).  No CPU fallback: every op raises if the
library is missing or a tensor is not on a CUDA device."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# AMQB_LIB: load another in-tree build of the same library (compile-time A/B variants made by tools/build_variant.sh)
LIB_PATH = os.environ.get("AMQB_LIB") or os.path.join(_HERE, "lib", "libamqb.so")
_lib: Optional[ctypes.CDLL] = None

LAYOUT_HQQ, LAYOUT_GPTQ, LAYOUT_FT, LAYOUT_NATIVE = 0, 1, 2, 3
PRO_NONE, PRO_RMSNORM, PRO_SILU_MUL, PRO_MUL = 0, 1, 2, 3


class ArCtx(ctypes.Structure):
    """amqb_ar_ctx (include/amqb.h)."""
    _fields_ = [("peer_bufs", ctypes.c_void_p * 16), ("rank", ctypes.c_int), ("world", ctypes.c_int),
                ("max_elems", ctypes.c_int), ("pos_dev", ctypes.c_void_p), ("gen_dev", ctypes.c_void_p)]


class GemvProblem(ctypes.Structure):
    """amqb_gemv_problem (include/amqb.h)."""
    _fields_ = [
        ("bits", ctypes.c_int), ("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
        ("w_native", ctypes.c_void_p), ("x", ctypes.c_void_p), ("ldx", ctypes.c_int),
        ("y", ctypes.c_void_p), ("ldy", ctypes.c_int),
        ("bias", ctypes.c_void_p), ("residual", ctypes.c_void_p),
        ("prologue", ctypes.c_int), ("gamma", ctypes.c_void_p), ("eps", ctypes.c_float),
        ("allreduce", ctypes.POINTER(ArCtx)), ("ar_call", ctypes.c_int), ("act", ctypes.c_int), ("after_gemv", ctypes.c_int),
    ]


class GemmProblem(ctypes.Structure):
    """amqb_gemm_problem (include/amqb.h)."""
    _fields_ = [("bits", ctypes.c_int), ("N", ctypes.c_int), ("w_native", ctypes.c_void_p), ("y", ctypes.c_void_p),
                ("bias", ctypes.c_void_p)]


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"amq_b200: {LIB_PATH} not found. Build it with `python -m amq_b200.build` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
        L = ctypes.CDLL(LIB_PATH)
        L.amqb_last_error_string.restype = ctypes.c_char_p
        for name in ("amqb_native_bytes", "amqb_workspace_bytes", "amqb_gemm_workspace_bytes",
                     "amqb_hqq_quantize_workspace_bytes", "amqb_ar_buffer_bytes", "amqb_ar_rows_buffer_bytes",
                     "amqb_attn_split_workspace_bytes"):
            if hasattr(L, name):
                getattr(L, name).restype = ctypes.c_size_t
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().amqb_last_error_string().decode()
        raise RuntimeError(f"amqb {what} failed (status {rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("amq_b200: tensor is not on a CUDA device; this path has no CPU fallback")
    if not t.is_contiguous():
        raise RuntimeError("amq_b200: tensor must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def cur_stream() -> ctypes.c_void_p:
    # torch.cuda.current_stream() costs ~14 us of Python per call (measured, tools/profile_module_path.py); the raw
    # handle is what the C ABI wants anyway
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
