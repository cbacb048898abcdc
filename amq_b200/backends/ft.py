"""FT_QuantLinear — drop-in for /root/reference/amq/kernel/hqq/hqq/backends/ft.py:57-145 (the
4-bit AWQ / TRT-LLM interleaved layout module) plus `pack_intweight` (:15-55) and the HQQ -> FT
converters (:148-205).  Same constructor, buffers (`qweight int16 [N/4, K]`, `scales`,
`scaled_zeros` fp16 [K/G, N], optional `bias`) and `pack`, so `*_FTLinear.pt` files load
unchanged.  `forward` serves every row count with the sm_100a kernels (the reference switches
gemv_4bit -> gemm_4bit at 8 rows, ft.py:128-142)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .._lib import LAYOUT_FT


def pack_intweight(unpacked_qweight, interleave=4, kstride=64):
    """ft.py:15-55 on the GPU: int codes [N, K] -> int16 [N/4, K]."""
    if interleave != 4 or kstride != 64:
        raise NotImplementedError("pack_intweight: only interleave=4, kstride=64 (the values the reference uses, ft.py:119)")
    if not unpacked_qweight.is_cuda:
        raise RuntimeError("amq_b200.pack_intweight: CUDA tensor required (no CPU path)")
    N, K = unpacked_qweight.shape
    # codes -> (W = q, scale = 1, zero = 0) reuses the fused quantise+pack kernel exactly
    W = unpacked_qweight.to(torch.float16)
    G = K
    ones = torch.ones((N, 1), dtype=torch.float16, device=W.device)
    q, _, _ = ops.ft_pack(W, ones, torch.zeros_like(ones), G)
    return q


class FT_QuantLinear(nn.Module):
    def __init__(self, bits, infeatures, outfeatures, bias, dtype, group_size, name):
        super().__init__()
        assert bits in [4], "Only 4 bits is supported."
        assert dtype == torch.float16, "Only fp16 is supported."
        self.bits = bits
        self.infeatures = infeatures
        self.outfeatures = outfeatures
        self.group_size = group_size if group_size != -1 else infeatures
        self.interleave = 4
        assert infeatures % self.group_size == 0
        assert outfeatures % (32 // self.bits) == 0
        int16_pack_num = 16 // self.bits
        self.register_buffer("qweight", torch.empty((outfeatures // self.interleave,
                                                     infeatures // int16_pack_num * self.interleave), dtype=torch.int16))
        numgroup = infeatures // self.group_size if self.group_size > 0 else 1
        self.register_buffer("scales", torch.empty((numgroup, outfeatures), dtype=dtype))
        self.register_buffer("scaled_zeros", torch.empty((numgroup, outfeatures), dtype=dtype))
        if bias:
            self.register_buffer("bias", torch.empty((outfeatures), dtype=torch.float16))
        else:
            self.bias = None
        self.dtype = dtype
        self.name = name
        self.register_buffer("w_native", None, persistent=False)
        self._native_ok = None

    def post_init(self):
        if not self.qweight.is_cuda:
            raise RuntimeError("amq_b200.FT_QuantLinear: move the module to a CUDA device first (no CPU path)")
        N, K, G = self.outfeatures, self.infeatures, self.group_size
        if not ops.native_supported(4, N, K, G):
            raise RuntimeError(f"amq_b200.FT_QuantLinear: shape N={N} K={K} G={G} unsupported "
                               "(needs N % 32 == 0, K % 128 == 0, group 128)")
        self.w_native = ops.repack_ft(self.qweight, self.scales, self.scaled_zeros, N, K, G)
        self._native_ok = True
        return self

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.w_native = None
        self._native_ok = None

    def pack(self, weight, scales, zeros, sym: bool = False):
        """ft.py:103-126."""
        if not weight.is_cuda:
            raise RuntimeError("amq_b200.FT_QuantLinear.pack: CUDA tensors required (no CPU path)")
        self.sym = sym
        if sym:
            zeros = zeros + 2 ** (self.bits - 1)
        q, s, sz = ops.ft_pack(weight.data, scales.to(weight.device), zeros.to(weight.device), self.group_size)
        self.qweight, self.scales, self.scaled_zeros = q, s, sz
        self.w_native = None
        self._native_ok = None

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("amq_b200.FT_QuantLinear.forward: CUDA tensor required (no CPU fallback)")
        if self._native_ok is None:
            self.post_init()
        out_shape = x.shape[:-1] + (self.outfeatures,)
        x2 = x.reshape(-1, x.shape[-1])
        if x2.dtype != torch.float16:
            x2 = x2.half()
        y = ops.linear_forward(4, self.w_native, x2, self.outfeatures, self.infeatures, self.bias)
        return y.reshape(out_shape)

    forward_normal = forward

    def unpack_codes(self):
        """int codes [N, K] (the bit-exactness probe, amqb_unpack_codes)."""
        return ops.unpack_codes(self.qweight, 4, LAYOUT_FT, self.outfeatures, self.infeatures, self.group_size)


def patch_hqq_to_ft(layer, patch_params, load=False):
    """ft.py:148-201."""
    from ..core.quantize import HQQLinear, Quantizer
    if type(layer) is not HQQLinear:
        return layer
    hqq_layer = layer
    device = hqq_layer.device
    nbits = hqq_layer.meta["nbits"]
    group_size = hqq_layer.meta["group_size"]
    outfeatures, infeatures = hqq_layer.meta["shape"]
    bias = hqq_layer.bias
    ft_layer = FT_QuantLinear(nbits, infeatures, outfeatures, bias is not None, torch.float16, group_size,
                              hqq_layer.name).to(device)
    if bias is not None and not load:
        ft_layer.bias = bias.detach().to(device=device, dtype=torch.float16).clone()
    if not load:
        W_deq = Quantizer.dequantize(hqq_layer.W_q, hqq_layer.meta)
        scales = hqq_layer.meta["scale"].reshape(outfeatures, -1)
        zeros = hqq_layer.meta["zero"].reshape(outfeatures, -1)
        ft_layer.pack(W_deq, scales, zeros, False)
    del hqq_layer.W_q
    del hqq_layer.meta
    del hqq_layer.bias
    return ft_layer


def patch_hqq_to_ft_load(layer, patch_params):
    return patch_hqq_to_ft(layer, patch_params, load=True)
