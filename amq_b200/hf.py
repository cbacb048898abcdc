"""AutoHQQHFModel — the proxy-checkpoint interface AMQ's benchmark and search start from
(/root/reference/amq/kernel/hqq/hqq/models/base.py:240-250 setup_model, :267-401 quantize_model, :405-434 serialize /
save_quantized, :464-543 from_quantized; hf/base.py:8-44 config caching / model creation; called first thing by
amq/amq_speed_benchmark.py:129-131 and amq/amq_quantization_proxy.py).

Same on-disk format: `<save_dir>/config.json` (the HF config, `architectures` set) + `<save_dir>/qmodel.pt` =
torch.save({module_name: state_dict}) with the non-encoded HQQLinear state dict for every quantized
linear and the plain tensors of every other leaf module.  Checkpoints written by the reference load here and vice versa.
The quantize / dequantize / forward arithmetic behind HQQLinear is this library's CUDA path; everything in this file is
module-tree bookkeeping."""
from __future__ import annotations

import contextlib
import os
from typing import Dict, Iterator, Optional, Tuple, Union

import torch
from torch import nn

from .core.quantize import HQQLinear

_IGNORE_LINEAR = ["lm_head"]          # models/base.py:43


def name_to_linear_tag(name: str) -> str:
    """model.layers.31.self_attn.k_proj -> self_attn.k_proj (models/base.py:61-69)."""
    return ".".join(n for n in name.split(".") if n not in ("model", "layers") and not n.isnumeric())


def _parent_of(model: nn.Module, name: str) -> nn.Module:
    parent = model
    for part in name.split(".")[:-1]:
        parent = parent._modules[part]
    return parent


def _leaves(model: nn.Module) -> Iterator[Tuple[str, nn.Module]]:
    for name, module in model.named_modules():
        if name and len(module._modules) == 0:
            yield name, module


def _is_quant_linear(name: str, module: nn.Module) -> bool:
    return type(module) in (nn.Linear, HQQLinear) and name.split(".")[-1] not in _IGNORE_LINEAR


@contextlib.contextmanager
def _empty_parameters():
    """Parameters land on the meta device while the model is built (buffers such as the rotary tables stay real): what
    accelerate's init_empty_weights does for the reference's create_model (hf/base.py:38-39)."""
    register = nn.Module.register_parameter

    def meta_register(module, name, param):
        register(module, name, param)
        if param is not None and module._parameters[name] is not None:
            p = module._parameters[name]
            module._parameters[name] = nn.Parameter(p.detach().to("meta"), requires_grad=p.requires_grad)

    nn.Module.register_parameter = meta_register
    try:
        yield
    finally:
        nn.Module.register_parameter = register


class AutoHQQHFModel:
    # ---- naming (models/base.py:165-214, 240-250)
    @classmethod
    def autoname_modules(cls, model: nn.Module) -> None:
        for name, module in model.named_modules():
            module.name = name

    @classmethod
    def set_auto_linear_tags(cls, model: nn.Module, ignore=_IGNORE_LINEAR) -> None:
        if not hasattr(model, "linear_tags"):
            tags = []
            for name, module in model.named_modules():
                if type(module) in (nn.Linear, HQQLinear) and name.split(".")[-1] not in ignore:
                    tag = name_to_linear_tag(name)
                    if tag not in tags:
                        tags.append(tag)
            model.linear_tags = tags
            model.base_class = cls

    @classmethod
    def setup_model(cls, model: nn.Module) -> nn.Module:
        cls.autoname_modules(model)
        cls.set_auto_linear_tags(model)
        return model

    @classmethod
    def get_config_file(cls, save_dir: str) -> str:
        return os.path.join(save_dir, "config.json")

    @classmethod
    def get_weight_file(cls, save_dir: str) -> str:
        return os.path.join(save_dir, "qmodel.pt")

    # ---- quantize (models/base.py:267-401)
    @classmethod
    def quantize_model(cls, model: nn.Module, quant_config: dict, compute_dtype: torch.dtype = torch.float16,
                       device: Union[str, torch.device] = "cuda") -> None:
        """Every nn.Linear outside `lm_head` becomes an HQQLinear, every other leaf module moves to `device` in
        `compute_dtype`.  quant_config: one config for all linears, or {linear_tag: config or None} (None: that tag stays
        an nn.Linear), as in the reference.  One device (AMQ dispatches proxies across GPUs afterwards,
        amq/utils/dispatch.py, by moving whole modules)."""
        if getattr(model, "hqq_quantized", False):
            print("Model was already quantized")
            return
        if not isinstance(device, (str, torch.device)):
            raise NotImplementedError("amq_b200.AutoHQQHFModel.quantize_model: a single device (str / torch.device)")
        cls.setup_model(model)
        per_tag = any(k in model.linear_tags for k in quant_config.keys())
        params: Dict[str, Optional[dict]] = {k: (quant_config.get(k) if per_tag else quant_config) for k in model.linear_tags}
        model.eval()
        for p in model.parameters():
            p.requires_grad = False
        for name, module in list(_leaves(model)):
            if _is_quant_linear(name, module):
                cfg = params.get(name_to_linear_tag(name))
                if type(module) is HQQLinear:
                    continue
                if cfg is not None:
                    new = HQQLinear(module, cfg, compute_dtype=compute_dtype, device=device)
                else:
                    new = module.to(device=device, dtype=compute_dtype)
            else:
                new = module.to(device=device, dtype=compute_dtype)
            new.name = name
            new.device = device
            setattr(_parent_of(model, name), name.split(".")[-1], new)
        model.base_class = cls
        model.hqq_quantized = True

    # ---- save (models/base.py:405-434, hf/base.py:10-16)
    @classmethod
    def serialize_weights(cls, model: nn.Module, verbose: bool = False) -> dict:
        weights = {}
        for name, module in _leaves(model):
            module.encoded_state_dict = False            # plain Python values, not the safetensors encoding
            sd = module.state_dict()
            if len(sd) > 0:
                weights[name] = dict(sd)
        return weights

    @classmethod
    def save_quantized(cls, model: nn.Module, save_dir: str, verbose: bool = False) -> None:
        os.makedirs(save_dir, exist_ok=True)
        model.config.architectures = [model.__class__.__name__]
        model.config.save_pretrained(save_dir)
        torch.save(cls.serialize_weights(model, verbose=verbose), cls.get_weight_file(save_dir))

    # ---- load (models/base.py:464-543, hf/base.py:19-41)
    @classmethod
    def create_model(cls, save_dir: str, kwargs: dict) -> nn.Module:
        import transformers
        model_kwargs = {k: kwargs[k] for k in ("attn_implementation",) if k in kwargs}
        config = transformers.AutoConfig.from_pretrained(cls.get_config_file(save_dir))
        archs = config.architectures or []
        auto = transformers.AutoModel
        if len(archs) == 1 and "CausalLM" in archs[0]:
            auto = transformers.AutoModelForCausalLM
        elif len(archs) == 1 and "SequenceClassification" in archs[0]:
            auto = transformers.AutoModelForSequenceClassification
        with _empty_parameters():
            model = auto.from_config(config, **model_kwargs)
        return model

    @classmethod
    def from_quantized(cls, save_dir: str, compute_dtype: torch.dtype = torch.float16, device: Union[str, torch.device] = "cuda",
                       cache_dir: Optional[str] = "", adapter: Optional[str] = None, **kwargs) -> nn.Module:
        if adapter is not None:
            raise NotImplementedError("amq_b200.AutoHQQHFModel.from_quantized: LoRA adapters are not part of AMQ's path")
        save_dir = os.path.join(cache_dir, save_dir) if cache_dir else save_dir
        if not os.path.exists(cls.get_weight_file(save_dir)):
            raise Exception("Weight file missing. Check your cache directory.")
        if not os.path.exists(cls.get_config_file(save_dir)):
            raise Exception("Config file missing. Check your cache directory.")
        model = cls.create_model(save_dir, kwargs)
        model.save_dir = save_dir
        cls.setup_model(model)
        weights = torch.load(cls.get_weight_file(save_dir), map_location=device, weights_only=True)
        model.eval()
        with torch.no_grad():
            for name, module in list(_leaves(model)):
                if name not in weights:
                    new = module.to(device=device, dtype=compute_dtype)       # no parameters of its own (rotary tables ...)
                else:
                    sd = weights[name]
                    if "W_q" in sd:
                        new = HQQLinear(None, None, compute_dtype=compute_dtype, device=device)
                        new.load_state_dict(sd)
                    else:
                        new = module
                        for key, t in sd.items():
                            setattr(new, key, nn.Parameter(t.to(device=device, dtype=compute_dtype), requires_grad=False))
                new.name = name
                setattr(_parent_of(model, name), name.split(".")[-1], new)
        # weight tying done by HF at init is lost on the meta device: the checkpoint stores both tensors
        model.hqq_quantized = True
        model.base_class = cls
        return model


class GraphedHFDecoder:
    """Greedy token-by-token decoding of an HF causal LM THROUGH ITS OWN forward — the module path of the reference's
    benchmark (amq/amq_speed_benchmark.py:231-251: linears swapped for GPTQLinear / FT_QuantLinear by setattr,
    amq/utils/speed.py:23-46: generate) — with the single-token forward captured once in a CUDA graph over a static KV
    cache.  The arithmetic is HF's module tree calling this library's module forwards; the graph only removes the host
    work between the ~1500 launches of a token (HF eager: ~17 ms of Python per token on the 7B shape; the graph: GPU time).

        dec = GraphedHFDecoder(model, max_cache_len=256)
        tokens = dec.generate(input_ids, max_new_tokens=128)          # [B, prompt + new], greedy

    The next-token argmax, its feedback into the input buffer and the position advance are part of the captured step, so a
    replay needs no host input at all."""

    def __init__(self, model: nn.Module, max_cache_len: int):
        from transformers import StaticCache
        self.model = model
        self.max_cache_len = int(max_cache_len)
        self.device = next(p for p in model.parameters()).device
        self._StaticCache = StaticCache
        self.cache = None
        self.ids = None
        self.pos = torch.zeros(1, dtype=torch.long, device=self.device)
        self.graph = None
        self.logits = None
        self._pos_h = None             # host mirror of the position; None until a prompt has been consumed

    def _forward_step(self):
        out = self.model(self.ids, past_key_values=self.cache, cache_position=self.pos, use_cache=True)
        self.logits = out.logits
        self.ids.copy_(out.logits[:, -1:].argmax(-1))
        self.pos.add_(1)

    @torch.inference_mode()
    def prefill(self, input_ids: torch.Tensor) -> None:
        """Consume the prompt eagerly (one forward over all prompt tokens), leave the first generated token in self.ids."""
        B, T = input_ids.shape
        if T + 1 > self.max_cache_len:
            raise ValueError("GraphedHFDecoder: prompt does not fit max_cache_len")
        if self.cache is None or self.ids is None or self.ids.shape[0] != B:
            self.cache = self._StaticCache(config=self.model.config, max_cache_len=self.max_cache_len)
            self.ids = torch.zeros(B, 1, dtype=torch.long, device=self.device)
            self.graph = None
        else:
            self.cache.reset()
        out = self.model(input_ids.to(self.device), past_key_values=self.cache,
                         cache_position=torch.arange(T, device=self.device), use_cache=True)
        self.ids.copy_(out.logits[:, -1:].argmax(-1))
        self.pos.fill_(T)
        self._pos_h = T

    @torch.inference_mode()
    def _capture(self) -> None:
        ids0, pos0 = self.ids.clone(), self.pos.clone()
        # a static cache layer keeps its own device-side fill count (transformers 5.x StaticLayer.cumulative_length, advanced
        # in place by every update): the warm-up steps must leave it where it was
        counters = [l.cumulative_length for l in getattr(self.cache, "layers", []) if torch.is_tensor(getattr(l, "cumulative_length", None))]
        saved = [c.clone() for c in counters]

        def rewind():
            self.ids.copy_(ids0); self.pos.copy_(pos0)
            for c, v in zip(counters, saved):
                c.copy_(v)
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(2):                       # warm-up at the current position (writes the cache row a real step rewrites)
                rewind()
                self._forward_step()
            rewind()
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                self._forward_step()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self.graph = g

    @torch.inference_mode()
    def step(self) -> torch.Tensor:
        """One token: replays the captured forward; returns the device buffer holding the NEW token ids [B, 1]."""
        if self._pos_h is None:
            raise RuntimeError("GraphedHFDecoder.step: call prefill(input_ids) first")
        if self._pos_h + 1 > self.max_cache_len:
            raise RuntimeError("GraphedHFDecoder: static cache is full")
        if self.graph is None:
            self._capture()
        self.graph.replay()
        self._pos_h += 1
        return self.ids

    @torch.inference_mode()
    def generate(self, input_ids: torch.Tensor, max_new_tokens: int) -> torch.Tensor:
        self.prefill(input_ids)
        out = [input_ids.to(self.device), self.ids.clone()]
        for _ in range(max_new_tokens - 1):
            out.append(self.step().clone())
        return torch.cat(out, dim=1)
