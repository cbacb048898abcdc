"""BitPack — HQQ's bit packing (/root/reference/amq/kernel/hqq/hqq/core/bitpack.py:24-110) as CUDA
kernels (amqb_hqq_pack / amqb_hqq_unpack).  Same static-method names and tensor layouts: packing is
block-strided over dim 0.  CUDA tensors only."""
from __future__ import annotations

import torch
from torch import Tensor, uint8

from .. import ops


class BitPack:
    @staticmethod
    def pack_4bit_u8(W_q: Tensor) -> Tensor:
        return ops.hqq_pack(4, W_q)

    @staticmethod
    def unpack_4bit_u8(W_q: Tensor, dtype=uint8) -> Tensor:
        return ops.hqq_unpack(4, W_q).to(dtype)

    @staticmethod
    def pack_2bit_u8(W_q: Tensor) -> Tensor:
        return ops.hqq_pack(2, W_q)

    @staticmethod
    def unpack_2bit_u8(W_q: Tensor, dtype=uint8) -> Tensor:
        return ops.hqq_unpack(2, W_q).to(dtype)

    @staticmethod
    def pack_3bit_32(W_q_in: Tensor) -> Tensor:
        return ops.hqq_pack(3, W_q_in)

    @staticmethod
    def unpack_3bit_32(W_q: Tensor, dtype=uint8) -> Tensor:
        return ops.hqq_unpack(3, W_q).to(dtype)
