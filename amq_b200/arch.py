"""Per-layer bit-config handling: the search-output JSON the speed benchmark consumes
(/root/reference/amq/amq_speed_benchmark.py:209-229, produced by amq/search/optimizer.py:164-171),
the bits accounting (amq/utils/func.py:101-114) and a synthetic-arch sampler that draws per-linear
bits the way the search space does (amq/search/space.py:34-84)."""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

LINEARS = ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
           "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"]


@dataclass
class ModelShape:
    """Shape tables of amq/configs/{llama,mistral,qwen2}.json plus what a decode step needs."""
    name: str
    hidden: int
    inter: int
    n_heads: int
    n_kv_heads: int
    n_block: int
    vocab: int
    head_dim: int = 128
    rope_theta: float = 10000.0
    rms_eps: float = 1e-5
    qkv_bias: bool = False
    model_numel: int = 0

    @property
    def linear_shape(self) -> Dict[str, List[int]]:
        q = self.n_heads * self.head_dim
        kv = self.n_kv_heads * self.head_dim
        return {"self_attn.q_proj": [q, self.hidden], "self_attn.k_proj": [kv, self.hidden],
                "self_attn.v_proj": [kv, self.hidden], "self_attn.o_proj": [self.hidden, q],
                "mlp.gate_proj": [self.inter, self.hidden], "mlp.up_proj": [self.inter, self.hidden],
                "mlp.down_proj": [self.hidden, self.inter]}

    def config(self) -> Dict:
        return {"n_block": self.n_block, "linear": LINEARS, "linear_shape": self.linear_shape,
                "model_numel": self.model_numel}


MODELS = {
    # amq/configs/llama.json:2-27
    "Llama-2-7b-hf": ModelShape("Llama-2-7b-hf", 4096, 11008, 32, 32, 32, 32000, model_numel=6476005376),
    # amq/configs/llama.json:56-81
    "Llama-2-70b-hf": ModelShape("Llama-2-70b-hf", 8192, 28672, 64, 8, 80, 32000, model_numel=68451041280),
    # amq/configs/mistral.json:2-27
    "Mistral-7B-v0.3": ModelShape("Mistral-7B-v0.3", 4096, 14336, 32, 8, 32, 32768, rope_theta=1e6,
                                  model_numel=6979321856),
    # amq/configs/qwen2.json:2-27 (Qwen2.5-7B has Qwen2-7B's linear shapes)
    "Qwen2.5-7B": ModelShape("Qwen2.5-7B", 3584, 18944, 28, 4, 28, 152064, rope_theta=1e6, rms_eps=1e-6,
                             qkv_bias=True, model_numel=6525288448),
}


def get_bits_usage(arch: Dict, config: Dict, group_size: int = 128) -> float:
    """amq/utils/func.py:101-114: sum out*in*(bit + 32/G) / model_numel."""
    memory_usage = 0.0
    for linear_group, bits in arch["linear"].items():
        out_dim, in_dim = config["linear_shape"][linear_group]
        g = in_dim if group_size == -1 else group_size
        for bit in bits:
            eff = bit + (32 / g if bit < 16 else 0)
            memory_usage += int(out_dim) * int(in_dim) * eff
    return memory_usage / config["model_numel"]


def select_arch(stats: Dict, target_bits: float) -> Dict[str, List[int]]:
    """amq_speed_benchmark.py:209-229: entries within 0.05 bits of the target, most 4-bit linears wins."""
    archs = stats["archive"] + stats["candidates"]
    candidates = [a for a in archs if abs(a[-1] - target_bits) < 0.05]
    if not candidates:
        raise ValueError(f"no architecture within 0.05 bits of {target_bits}")
    bits = [np.concatenate([np.asarray(b) for b in a[0]["linear"].values()]) for a in candidates]
    count_4bit = [(b == 4.0).sum() for b in bits]
    return candidates[int(np.argmax(count_4bit))][0]["linear"]


def load_arch(path: Optional[str], target_bits: float, n_block: int) -> Dict[str, List[int]]:
    if path is not None:
        if not os.path.exists(path):
            raise FileNotFoundError(f"Arch file {path} not found")
        with open(path) as f:
            return select_arch(json.load(f), target_bits)
    assert target_bits in [2, 3, 4], "target bits should be 2, 3, 4 if arch_path is not provided"
    return {name: [int(target_bits)] * n_block for name in LINEARS}


def sample_arch(shape: ModelShape, target_bits: float, seed: int = 0, tol: float = 0.05,
                bits_range=(2, 3, 4), group_size: int = 128) -> Dict[str, List[int]]:
    """Synthetic search output: per-linear bits drawn like SearchSpace.sample (space.py:34-84) with
    np.random.seed(seed) until get_bits_usage is within `tol` of the target."""
    rs = np.random.RandomState(seed)
    cfg = shape.config()
    bits_range = list(bits_range)
    for _ in range(100000):
        prob = rs.rand(len(bits_range))
        p = prob / prob.sum()
        arch = {"linear": {name: rs.choice(bits_range, size=shape.n_block, p=p, replace=True).tolist()
                           for name in LINEARS}}
        if abs(get_bits_usage(arch, cfg, group_size) - target_bits) < tol:
            return arch["linear"]
    raise RuntimeError("could not sample an architecture at the requested bits")


def make_stats_file(path: str, shape: ModelShape, target_bits: float, n: int = 8, seed: int = 0) -> None:
    """Write an iter_N.stats-shaped JSON (optimizer.py:164-171) holding synthetic candidates."""
    cfg = shape.config()
    entries = []
    for i in range(n):
        lin = sample_arch(shape, target_bits, seed=seed + i)
        entries.append([{"linear": lin}, 0.0, get_bits_usage({"linear": lin}, cfg)])
    with open(path, "w") as f:
        json.dump({"archive": entries[: n // 2], "candidates": entries[n // 2:], "iteration": 0}, f)
