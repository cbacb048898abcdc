"""prepare_for_inference — /root/reference/amq/kernel/hqq/hqq/utils/patching.py:39-49, 143-223,
gptq / ft branches (the only backends AMQ selects, amq_speed_benchmark.py:137-139): walk the
module tree, replace every HQQLinear with the kernel module, cache the repacked state dict under
`load_path` on the first run and load it on later runs, then add the dummy `.weight` parameter HF
code expects."""
from __future__ import annotations

import os

import torch

from ..backends.autogptq import patch_hqq_to_gptq, patch_hqq_to_gptq_load
from ..backends.ft import patch_hqq_to_ft, patch_hqq_to_ft_load
from ..core.quantize import HQQLinear


def patch_linearlayers(model, fct, patch_param=None, verbose=False):
    def _patch_linear(m):
        for name, layer in m.named_children():
            if isinstance(layer, HQQLinear):
                layer.name = name
                setattr(m, name, fct(layer, patch_param))
            else:
                _patch_linear(layer)
    _patch_linear(model)


def patch_add_weight_param(layer, patch_param):
    """patching.py:76-91."""
    if hasattr(layer, "weight") is False:
        bufs = [b for b in layer.buffers()]
        device_ = bufs[0].device if bufs else patch_param["device"]
        layer.weight = torch.nn.Parameter(torch.zeros((1,), device=device_, dtype=patch_param["dtype"]),
                                          requires_grad=False)
    return layer


def _add_weight_params(model, patch_param):
    from ..backends.autogptq import GPTQLinear
    from ..backends.ft import FT_QuantLinear
    for m in model.modules():
        if isinstance(m, (GPTQLinear, FT_QuantLinear)):
            patch_add_weight_param(m, patch_param)


def prepare_for_inference(model, allow_merge=False, backend="default", verbose=False, load_path=None):
    if backend not in ("gptq", "ft"):
        raise NotImplementedError("amq_b200.prepare_for_inference: backend must be 'gptq' or 'ft' "
                                  "(the two AMQ uses, amq_speed_benchmark.py:137-139)")
    fresh, load = (patch_hqq_to_gptq, patch_hqq_to_gptq_load) if backend == "gptq" else (patch_hqq_to_ft, patch_hqq_to_ft_load)
    if load_path is not None and os.path.exists(load_path) is False:
        patch_linearlayers(model, fresh, verbose=verbose)
        print("Saving the model to", load_path)
        torch.save(model.state_dict(), load_path)
    elif load_path is not None and os.path.exists(load_path) is True:
        patch_linearlayers(model, load, verbose=verbose)
        print("Loading the model from", load_path)
        model.load_state_dict(torch.load(load_path, weights_only=True))
    else:
        print("No load_path provided, using the model as is")
        patch_linearlayers(model, fresh, verbose=verbose)
    try:
        dev, dt = next(model.parameters()).device, torch.float16
    except StopIteration:
        dev, dt = torch.device("cuda"), torch.float16
    _add_weight_params(model, {"device": dev, "dtype": dt})
    return model
