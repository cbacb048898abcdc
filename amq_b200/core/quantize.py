"""Quantizer / HQQLinear / BaseQuantizeConfig — the HQQ "quantization proxy" interface AMQ uses
(/root/reference/amq/kernel/hqq/hqq/core/quantize.py:35-199 Quantizer, :387-1073 HQQLinear,
:1076-1155 hqq_base_quant_config), restricted to what AMQ exercises: axis=1, channel-wise,
group-wise, bits in {2,3,4}, unquantised fp16 meta.  Same meta keys, `W_q` layout and state-dict
keys, so reference `qmodel.pt` entries load; the arithmetic runs in CUDA
kernels (amqb_hqq_quantize / amqb_hqq_dequant) instead of chains of torch ops."""
from __future__ import annotations

import copy
from typing import Union

import torch
from torch import Tensor, float16, int32, nn, uint8

from .. import ops
from .bitpack import BitPack


class Quantizer:
    SUPPORTED_BITS = [4, 3, 2]
    bit_to_packing = {4: "4bit_u8", 3: "3bit_32", 2: "2bit_u8"}
    packing_to_bit = {v: k for k, v in bit_to_packing.items()}
    pack = {"4bit_u8": BitPack.pack_4bit_u8, "3bit_32": BitPack.pack_3bit_32, "2bit_u8": BitPack.pack_2bit_u8}
    unpack = {"4bit_u8": BitPack.unpack_4bit_u8, "3bit_32": BitPack.unpack_3bit_32, "2bit_u8": BitPack.unpack_2bit_u8}
    unpack_view_dtype = {"4bit_u8": uint8, "3bit_32": int32, "2bit_u8": uint8}
    # Arithmetic of the half-quadratic solver.  The reference picks it from the DEVICE (optimize.py:231: fp16 on CUDA,
    # fp32 on the CPU).  None = like the reference on a GPU (fp16); torch.float32 = its CPU branch, which is the one the
    # oracle pins and this library reproduces bit for bit.  AMQB_HQQ_SOLVER=fp32|fp16 overrides.
    solver_dtype = None

    @classmethod
    def quantize(cls, tensor: Tensor, nbits: float = 4, channel_wise: bool = True, group_size: int = 64,
                 optimize: bool = True, round_zero: bool = False, axis: int = 0, bitpack: bool = True,
                 compute_dtype: Union[torch.dtype, None] = None, view_as_float: bool = False,
                 device: str = "cuda") -> tuple:
        """quantize.py:75-180.  Returns (W_q, meta) with meta['scale'] already inverted."""
        assert nbits in Quantizer.SUPPORTED_BITS, "nbits=" + str(nbits) + " not supported."
        if axis != 1 or not channel_wise or not optimize or view_as_float or group_size is None:
            raise NotImplementedError("amq_b200.Quantizer: AMQ's proxy setting only (axis=1, channel_wise, optimize, "
                                      "group-wise; amq_quantization_proxy.py:36)")
        if tensor.numel() % group_size != 0:
            raise AssertionError("group_size should be divisble by the total tensor dimensions. shape: "
                                 + str(tensor.shape) + ", group_size: " + str(group_size))
        W = tensor.to(device)
        if not W.is_cuda:
            raise RuntimeError("amq_b200.Quantizer.quantize: CUDA device required (no CPU path)")
        shape = W.shape
        import os
        sd = {"fp32": torch.float32, "fp16": torch.float16}.get(os.environ.get("AMQB_HQQ_SOLVER", ""), cls.solver_dtype)
        if sd is None:
            sd = torch.float16
        packable = bitpack and (nbits == 3 or (W.numel() // group_size) % (2 if nbits == 4 else 4) == 0)
        res = ops.hqq_quantize(W.reshape(shape[0], -1), int(nbits), group_size, round_zero, solver_dtype=sd,
                               packed=bool(packable), want_codes=not packable)
        codes, scale, zero = res[0], res[1], res[2]
        meta = {"nbits": nbits, "group_size": group_size, "shape": shape, "scale": scale, "zero": zero,
                "axis": axis, "packing": Quantizer.bit_to_packing[nbits]}
        meta["unpack_view_dtype"] = Quantizer.unpack_view_dtype[meta["packing"]]
        meta["view_as_float"] = view_as_float
        if packable:
            W_q = res[4]                                    # packed in the quantize pass itself
        elif bitpack:
            W_q = Quantizer.pack[meta["packing"]](codes)
        else:
            W_q = codes.to(tensor.dtype)
            meta["packing"] = None
        return W_q, meta

    @classmethod
    def dequantize(cls, W_q: Tensor, meta: dict) -> Tensor:
        """quantize.py:183-199: ((unpack(W_q) - zero) * scale).reshape(shape), fp16, fused."""
        if not meta.get("packing"):
            raise NotImplementedError("amq_b200.Quantizer.dequantize: packed W_q only")
        compute_dtype = meta["compute_dtype"] if ("compute_dtype" in meta) else float16
        if compute_dtype != float16:
            raise NotImplementedError("amq_b200.Quantizer.dequantize: fp16 compute dtype only (AMQ's setting)")
        N, K = meta["shape"]
        bits = Quantizer.packing_to_bit[meta["packing"]]
        return ops.hqq_dequant(bits, W_q, meta["scale"].reshape(-1), meta["zero"].reshape(-1), N, K, meta["group_size"])

    @classmethod
    def cuda(cls, W_q: Tensor, meta: dict, device) -> tuple:
        """quantize.py:202-217: tensors to device, float meta cast to compute_dtype."""
        compute_dtype = meta["compute_dtype"] if ("compute_dtype" in meta) else float16
        if W_q is not None:
            W_q = W_q.to(device).contiguous()
        for key in meta:
            if isinstance(meta[key], torch.Tensor):
                t = meta[key]
                meta[key] = (t.to(compute_dtype) if torch.is_floating_point(t) else t).to(device).contiguous()
        return W_q, meta


def hqq_base_quant_config(nbits: int = 4, group_size: int = 64, quant_zero: bool = False, quant_scale: bool = False,
                          offload_meta: bool = False, view_as_float: bool = False, axis: int = 1):
    """quantize.py:1076-1152 (meta quantisation / offload are deprecated there and unsupported here)."""
    assert nbits in Quantizer.SUPPORTED_BITS, "nbits value not supported. Check Quantizer.SUPPORTED_BITS."
    if group_size is not None:
        assert group_size % 8 == 0, "Invalid group_size param: the value should be a multiple of 8."
    if quant_zero or quant_scale or offload_meta:
        raise NotImplementedError("amq_b200: quantised / offloaded meta-data is not part of AMQ's path")
    return {
        "weight_quant_params": {"nbits": nbits, "channel_wise": True, "group_size": group_size, "optimize": True,
                                "round_zero": True if nbits == 4 else False, "axis": axis,
                                "view_as_float": view_as_float},
        "scale_quant_params": None,
        "zero_quant_params": None,
        "offload_meta": offload_meta,
    }


BaseQuantizeConfig = hqq_base_quant_config

_META_KEYS = ["nbits", "group_size", "shape", "scale", "zero", "axis", "packing", "unpack_view_dtype",
              "view_as_float", "quant_scale", "quant_zero", "compute_dtype"]
_CFG_KEYS = ["nbits", "channel_wise", "group_size", "optimize", "round_zero", "axis", "view_as_float"]


def _decode(v, kind):
    """Inverse of the reference's safetensors encoding (core/utils.py:37-69)."""
    if not isinstance(v, torch.Tensor):
        return v
    if kind is torch.Size:
        return torch.Size(v.tolist())
    if kind is bool:
        return bool(v.item())
    if kind is int:
        return int(v.item())
    if kind is str:
        return "".join(chr(i) for i in v.tolist())
    if kind is torch.dtype:
        return getattr(torch, "".join(chr(i) for i in v.tolist()).replace("torch.", ""))
    return v


_KIND = {"nbits": int, "group_size": int, "shape": torch.Size, "axis": int, "packing": str,
         "unpack_view_dtype": torch.dtype, "view_as_float": bool, "quant_scale": bool, "quant_zero": bool,
         "compute_dtype": torch.dtype, "channel_wise": bool, "optimize": bool, "round_zero": bool}


class HQQLinear(nn.Module):
    """quantize.py:387-1073, inference subset: quantise an nn.Linear, dequantise, forward
    (x @ dequantize().T + bias), HQQ-format state dict."""

    def __init__(self, linear_layer: Union[nn.Module, None], quant_config: dict, del_orig: bool = True,
                 compute_dtype: torch.dtype = float16, device: str = "cuda", initialize: bool = True):
        super().__init__()
        self.ready = False
        self.bias = None
        self.device = device
        self.compute_dtype = compute_dtype
        self.quant_config = copy.deepcopy(quant_config)
        self.offload_meta = self.quant_config.pop("offload_meta", False) if self.quant_config is not None else None
        self.del_orig = del_orig
        self.linear_layer = linear_layer
        self.W_q = None
        self.meta = None
        self.name = None
        self.encoded_state_dict = False
        if initialize and linear_layer is not None:
            self.initialize()

    def is_initialized(self):
        return not (self.W_q is None or self.meta is None)

    def initialize(self):
        wq = self.quant_config["weight_quant_params"]
        if wq["group_size"] is None:
            wq["group_size"] = self.linear_layer.in_features
        self.quantize(self.linear_layer.weight.data, **self.quant_config)
        self.bias = (None if self.linear_layer.bias is None
                     else self.linear_layer.bias.clone().to(device=self.device, dtype=self.compute_dtype))
        if self.del_orig:
            del self.linear_layer
            self.linear_layer = None

    @classmethod
    def from_weights(cls, weight, bias, quant_config, compute_dtype=float16, device="cuda", del_orig=True):
        dummy = torch.nn.Linear(1, 1, bias=False)
        dummy.in_features, dummy.out_features = weight.shape[1], weight.shape[0]
        dummy.weight.data = weight
        dummy.bias = bias
        return cls(dummy, quant_config=quant_config, compute_dtype=compute_dtype, device=device, del_orig=del_orig)

    def quantize(self, W: Tensor, weight_quant_params: dict, scale_quant_params=None, zero_quant_params=None) -> None:
        self.in_features, self.out_features = W.t().shape
        W_q, meta = Quantizer.quantize(W, device=self.device, compute_dtype=self.compute_dtype, **weight_quant_params)
        meta.update({"quant_scale": False, "quant_zero": False})
        self.W_q, self.meta = W_q, meta
        self.cuda(self.device)
        self.ready = True

    def cuda(self, device):
        self.meta["compute_dtype"] = self.compute_dtype
        data = self.W_q.data if isinstance(self.W_q, nn.Parameter) else self.W_q
        data, self.meta = Quantizer.cuda(data, self.meta, device)
        self.W_q = nn.Parameter(data, requires_grad=False)
        if self.bias is not None:
            self.bias = self.bias.to(device=device, dtype=self.compute_dtype)
        self.device = device
        return self

    def dequantize(self):
        assert self.ready, "model was not quantized"
        return Quantizer.dequantize(self.W_q.data, self.meta)

    def unpack_codes(self):
        N, K = self.meta["shape"]
        from .._lib import LAYOUT_HQQ
        return ops.unpack_codes(self.W_q.data, int(self.meta["nbits"]), LAYOUT_HQQ, N, K, self.meta["group_size"])

    def matmul(self, x: Tensor, transpose: bool = True) -> Tensor:
        """quantize.py:880-882.  x @ W_r.T through the fused kernels; the untransposed product only serves the
        reference's backprop variants, which are not part of AMQ's path."""
        if not transpose:
            raise NotImplementedError("amq_b200.HQQLinear.matmul: transpose=False (backprop) is not on AMQ's path")
        bias, self.bias = self.bias, None
        try:
            return self.forward(x)
        finally:
            self.bias = bias

    # ---- forward (quantize.py:880-898 `forward_pytorch`): x @ ((W_q - zero) * scale).T + bias.  ONE path: W_q is
    # transcoded once, integer-exactly, HQQ layout -> kernel-native records, and every forward is the
    # dequant-fused decode GEMV (M <= 16) or the tcgen05 prefill GEMM — no [N, K] fp16 weight is materialised per
    # call and there is no dequantise + library-GEMM branch.  `set_backend` is kept for call-site compatibility
    # (quantize.py:393-418); every reference backend name maps onto the fused kernels.
    backend = "fused"

    @classmethod
    def set_backend(cls, backend):
        name = getattr(backend, "value", backend)
        if name not in ("fused", "pytorch", "forward_pytorch", "forward_pytorch_backprop", "forward_pytorch_compile",
                        "forward_pytorch_backprop_compile", "forward_aten", "forward_aten_backprop"):
            raise ValueError(f"HQQLinear.set_backend: unknown backend {backend!r}")
        cls.backend = "fused"

    def native_weight(self):
        """Kernel-native copy of (W_q, scale, zero); None when the layer does not meet the native layout's
        preconditions (axis 1, group 128, N % 32 == 0, K % 128 == 0, 2/3/4 bits).  Rebuilt when W_q / meta are replaced."""
        key = (self.W_q.data_ptr(), self.meta["scale"].data_ptr(), self.meta["zero"].data_ptr())
        if getattr(self, "_native_key", None) == key:
            return self._w_native
        N, K = self.meta["shape"]
        bits, G = int(self.meta["nbits"]), self.meta["group_size"]
        self._native_key, self._w_native, self._gptq = key, None, None
        if self.meta.get("axis", 1) == 1 and self.meta["nbits"] == bits and ops.native_supported(bits, N, K, G) \
                and self.W_q.is_cuda:
            scale = self.meta["scale"].reshape(N, -1).to(torch.float16)
            zero = self.meta["zero"].reshape(N, -1).to(torch.float16)
            self._w_native = ops.pack_native(bits, self.unpack_codes(), scale, zero)
        return self._w_native

    def drop_native(self):
        """Free the kernel-native copy (rebuilt on the next forward).  The proxies keep W_q for state_dict() /
        dequantize(), so a layer that has run holds its weights twice: (bits + 0.25) / 8 bytes per weight more."""
        self._native_key, self._w_native, self._gptq = None, None, None

    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError("amq_b200.HQQLinear.forward: CUDA tensor required (no CPU fallback)")
        N, K = self.meta["shape"]
        bits = int(self.meta["nbits"])
        x2 = x.reshape(-1, K)
        if x2.dtype != torch.float16:
            x2 = x2.half()
        x2 = x2.contiguous()
        nat = self.native_weight()
        if nat is not None:
            out = ops.linear_forward(bits, nat, x2, N, K, self.bias)
        else:
            # shapes outside the native record grid (never AMQ's): the reference's own conversion (dequantise ->
            # GPTQLinear.pack, autogptq.py:318-325) once, then the any-shape GPTQ-layout kernel in 16-row slabs
            if self._gptq is None:
                G = self.meta["group_size"]
                self._gptq = ops.gptq_pack(bits, self.dequantize(), self.meta["scale"].reshape(N, -1),
                                           self.meta["zero"].reshape(N, -1), G)
            q, s, z = self._gptq
            outs = [ops.gemv_gptq_layout(bits, q, s, z, x2[i:i + 16].contiguous(), N, K, self.meta["group_size"], self.bias)
                    for i in range(0, x2.shape[0], 16)]
            out = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        return out.to(x.dtype).reshape(*x.shape[:-1], N)

    # ---- HQQ-format (de)serialisation, quantize.py:643-787
    def state_dict_keys(self):
        return set(["W_q", "bias", "offload_meta", "encoded_state_dict", "stores_quant_config"] + _META_KEYS + _CFG_KEYS)

    def state_dict(self, *args, **kwargs):
        if not self.is_initialized():
            return {k: None for k in self.state_dict_keys()}
        state = {"W_q": self.W_q}
        state.update(dict(self.meta))
        if self.bias is not None:
            state["bias"] = self.bias
        state["offload_meta"] = False
        state["stores_quant_config"] = True
        for k in self.quant_config["weight_quant_params"]:
            state[k] = self.quant_config["weight_quant_params"][k]
        if "destination" in kwargs and "prefix" in kwargs:
            for key, value in state.items():
                kwargs["destination"][kwargs["prefix"] + key] = value
        return state

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        """quantize.py:684-706: a parent model.load_state_dict() hands this layer the flat dict; gather the
        prefix + state_dict_keys() entries (removing them from the parent's dict, so strict loading does not report
        them as unexpected) and load them as a layer state dict."""
        layer_sd = {}
        for key in self.state_dict_keys():
            full = prefix + key
            if full in state_dict:
                layer_sd[key] = state_dict.pop(full)
            elif strict and key in ("W_q", "scale", "zero", "nbits", "group_size", "shape"):
                missing_keys.append(full)
        if "W_q" in layer_sd and layer_sd["W_q"] is not None:
            self.load_state_dict(layer_sd, strict=strict)

    def load_state_dict(self, state_dict, strict=True, assign=False):
        sd = dict(state_dict)
        sd.pop("encoded_state_dict", None)
        if sd.pop("stores_quant_config", False):
            self.quant_config = {"weight_quant_params": {k: _decode(sd[k], _KIND[k]) for k in _CFG_KEYS},
                                 "scale_quant_params": None, "zero_quant_params": None}
        W_q = sd.pop("W_q")
        self.bias = sd.pop("bias", None)
        sd.pop("offload_meta", None)
        self.meta = {k: _decode(v, _KIND.get(k)) for k, v in sd.items() if k in _META_KEYS}
        if "unpack_view_dtype" not in self.meta:
            self.meta["unpack_view_dtype"] = Quantizer.unpack_view_dtype[self.meta["packing"]]
        self.meta.setdefault("view_as_float", False)
        self.meta.setdefault("quant_scale", False)
        self.meta.setdefault("quant_zero", False)
        self.W_q = W_q.data if isinstance(W_q, nn.Parameter) else W_q
        self.cuda(self.device)
        self.ready = True
        self.in_features, self.out_features = self.meta["shape"][::-1]
