"""GPTQLinear — drop-in for the reference module of the same name
(/root/reference/amq/kernel/hqq/hqq/backends/autogptq.py:27-288) and its HQQ -> GPTQ converters
(:291-345).  Same constructor, same registered buffers (names / shapes / dtypes, so a reference
`*_GPTQLinear.pt` state dict loads unchanged), same `pack(W, scales, zeros)` and `forward(x)`.

What differs is underneath: `pack` is a CUDA kernel instead of a numpy row loop, `post_init()`
builds the kernel-native repack once (a non-persistent buffer, so state dicts stay
reference-identical), and `forward` calls the sm_100a kernels through the C ABI for every row
count (decode kernel for M <= 16, tensor-core kernel above) instead of switching to a torch
unpack+matmul at `kernel_switch_threshold`.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops


class GPTQLinear(nn.Module):
    QUANT_TYPE = "cuda-old"

    def __init__(self, bits, group_size, infeatures, outfeatures, bias, use_cuda_fp16=True,
                 kernel_switch_threshold=128, trainable=False, weight_dtype=torch.float16, outlierfeatures=32):
        super().__init__()
        if bits not in [2, 3, 4, 8]:
            raise NotImplementedError("Only 2,3,4,8 bits are supported.")
        if trainable:
            raise NotImplementedError("amq_b200.GPTQLinear is inference-only")
        self.infeatures = infeatures
        self.outfeatures = outfeatures
        self.bits = bits
        self.group_size = group_size if group_size != -1 else infeatures
        self.maxq = 2 ** self.bits - 1

        self.register_buffer("qweight", torch.zeros((infeatures // 32 * self.bits, outfeatures), dtype=torch.int32))
        self.register_buffer("zeros", torch.zeros((math.ceil(infeatures / self.group_size), outfeatures), dtype=torch.float))
        self.register_buffer("scales", torch.zeros((math.ceil(infeatures / self.group_size), outfeatures), dtype=torch.float))
        self.name = None
        if bias:
            self.register_buffer("bias", torch.zeros((outfeatures), dtype=weight_dtype))
        else:
            self.bias = None
        self.half_indim = self.infeatures // 2
        self.use_cuda_fp16 = use_cuda_fp16 if bits != 8 else False      # autogptq.py:86
        self.kernel_switch_threshold = kernel_switch_threshold   # kept for API parity; unused
        self.trainable = trainable
        # kernel-native repack (amq_b200/csrc/layout.cuh); rebuilt lazily after pack()/load_state_dict()
        self.register_buffer("w_native", None, persistent=False)
        self._native_ok = None

    # -- native repack ------------------------------------------------------------------------
    def post_init(self):
        """Build the kernel-native weight buffer from qweight/scales/zeros (no-op in the reference,
        autogptq.py:107-108).  Needs the buffers on a CUDA device."""
        if not self.qweight.is_cuda:
            raise RuntimeError("amq_b200.GPTQLinear: move the module to a CUDA device first (no CPU path)")
        N, K, G = self.outfeatures, self.infeatures, self.group_size
        ok = self.bits != 8 and ops.native_supported(self.bits, N, K, G)
        if ok:
            # scales / zeros hold fp16-exact values in the AMQ flow (they come from fp16 HQQ meta,
            # autogptq.py:112-114); the native layout stores them as fp16.  Otherwise keep fp32 meta
            # and use the GPTQ-layout kernel.
            ok = bool((self.scales.half().float() == self.scales).all() and
                      (self.zeros.half().float() == self.zeros).all())
        self._native_ok = ok
        self.w_native = ops.repack_gptq(self.bits, self.qweight, self.scales, self.zeros, N, K, G) if ok else None
        return self

    def _invalidate(self):
        self.w_native = None
        self._native_ok = None

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._invalidate()

    def pack(self, W, scales, zeros):
        """autogptq.py:111-156 on the GPU: W fp16 [N,K] (dequantised), scales/zeros [N, K/G]."""
        if not W.is_cuda:
            raise RuntimeError("amq_b200.GPTQLinear.pack: CUDA tensors required (no CPU path)")
        q, s, z = ops.gptq_pack(self.bits, W, scales.to(W.device), zeros.to(W.device), self.group_size)
        self.qweight, self.scales, self.zeros = q, s, z
        self._invalidate()

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("amq_b200.GPTQLinear.forward: CUDA tensor required (no CPU fallback)")
        x_dtype = x.dtype
        out_shape = x.shape[:-1] + (self.outfeatures,)
        x2 = x.reshape(-1, x.shape[-1])
        if x_dtype != torch.float16:
            x2 = x2.half()      # the reference kernels cast too (autogptq.py:165-169)
        if self.bits == 8 and x2.shape[0] < self.kernel_switch_threshold:
            # autogptq.py:86,204-205: 8-bit modules have use_cuda_fp16 = False and the reference's small-M branch
            # raises exactly this; only the large-M branch (:245-283) serves them
            raise NotImplementedError("Only use_cuda_fp16=True is supported.")
        if self._native_ok is None:
            self.post_init()
        N, K = self.outfeatures, self.infeatures
        if self._native_ok:
            out = ops.linear_forward(self.bits, self.w_native, x2, N, K, self.bias)
        else:
            outs = [ops.gemv_gptq_layout(self.bits, self.qweight, self.scales, self.zeros, x2[i:i + 16], N, K,
                                         self.group_size, self.bias) for i in range(0, x2.shape[0], 16)]
            out = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        return out.to(dtype=x_dtype).reshape(out_shape)


def patch_hqq_to_gptq(layer, patch_params, load=False):
    """autogptq.py:291-341: HQQLinear -> GPTQLinear (dequantise + pack, or an empty shell when the
    packed state dict is loaded afterwards)."""
    from ..core.quantize import HQQLinear, Quantizer
    if type(layer) is not HQQLinear:
        return layer
    hqq_layer = layer
    device = hqq_layer.device
    nbits = hqq_layer.meta["nbits"]
    group_size = hqq_layer.meta["group_size"]
    outfeatures, infeatures = hqq_layer.meta["shape"]
    bias = hqq_layer.bias
    gptq_layer = GPTQLinear(nbits, group_size, infeatures, outfeatures, bias is not None).to(device)
    gptq_layer.name = hqq_layer.name
    if bias is not None and not load:
        gptq_layer.bias = bias.detach().to(device=device, dtype=torch.float16).clone()
    if not load:
        W_deq = Quantizer.dequantize(hqq_layer.W_q, hqq_layer.meta)
        scales = hqq_layer.meta["scale"].reshape(outfeatures, -1)
        zeros = hqq_layer.meta["zero"].reshape(outfeatures, -1)
        gptq_layer.pack(W_deq, scales, zeros)
    del hqq_layer.W_q
    del hqq_layer.meta
    del hqq_layer.bias
    return gptq_layer


def patch_hqq_to_gptq_load(layer, patch_params):
    return patch_hqq_to_gptq(layer, patch_params, load=True)
