"""amq_b200 — B200-native (sm_100a) implementation of AMQ's quantized-linear hot path.

Host-side mirror of the reference's plugin interface for this path
(/root/reference/amq/kernel/hqq/hqq): same module constructors, buffer names / shapes / dtypes,
`pack`, `forward`, `prepare_for_inference`; the kernels live behind the C ABI in include/amqb.h.
"""
from . import _lib  # noqa: F401
from .backends.autogptq import GPTQLinear, patch_hqq_to_gptq, patch_hqq_to_gptq_load  # noqa: F401
from .backends.ft import FT_QuantLinear, pack_intweight, patch_hqq_to_ft, patch_hqq_to_ft_load  # noqa: F401
from .core.bitpack import BitPack  # noqa: F401
from .core.quantize import BaseQuantizeConfig, HQQLinear, Quantizer  # noqa: F401
from .utils.patching import prepare_for_inference  # noqa: F401
from .hf import AutoHQQHFModel  # noqa: F401

__version__ = "0.1.0"
