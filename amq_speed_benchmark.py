#!/usr/bin/env python
"""Speed benchmark of an AMQ mixed-precision decoder on the B200-native kernels.

Command-line mirror of the reference's /root/reference/amq/amq_speed_benchmark.py (flags :103-125, arch selection
:209-229, per-linear swap :231-251, result file :199-203, 290-293):

    python amq_speed_benchmark.py --model_name Llama-2-7b-hf --target_bits 3 --arch_path iter_200.stats \
        --tps --gemv --gemm --ttft --memory --peak_memory --file_name llama7b_3bit.json

What differs, and why: there are no checkpoints offline, so the decoder is random-init with the named
architecture's shapes (amq/configs/*.json) and every linear is created directly in the kernel-native packed layout
at the bit width the arch assigns to it (the reference swaps pre-quantized GPTQLinear / FT_QuantLinear modules into
an HF model; that module path is amq_b200.utils.patching + amq_b200.backends).  `--model_path`, `--save_path` and
`--use_ft` are accepted and ignored.  Without `--arch_path`, `--target_bits` must be 2, 3 or 4 (uniform), exactly as
the reference; `--synthetic_arch` draws a mixed arch at a fractional target the way the search space samples.
The fp16 base-model pass of the reference is not repeated (this package has no dense fp16 decoder).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser()
    parser.add_argument('--model_path', type=str, help='model path (ignored: random-init weights)', default='meta-llama')
    parser.add_argument('--model_name', type=str, help='model name', default='Llama-2-7b-hf')
    parser.add_argument('--save_path', type=str, help='save path (ignored: no checkpoints offline)', default='/SSD/hqq')
    parser.add_argument('--use_ft', action='store_true', help='accepted for compatibility (one attention path here)')

    parser.add_argument('--batch_size', type=int, help='batch size', default=1)
    parser.add_argument('--seq_length', type=int, help='sequence length', default=64)
    parser.add_argument('--gen_length', type=int, help='generation length', default=128)

    parser.add_argument('--tps', action='store_true', help='token per second')
    parser.add_argument('--gemm', action='store_true', help='gemm')
    parser.add_argument('--gemv', action='store_true', help='gemv')
    parser.add_argument('--ttft', action='store_true', help='ttft')
    parser.add_argument('--memory', action='store_true', help='memory')
    parser.add_argument('--peak_memory', action='store_true', help='peak memory & It only works with TPS')

    parser.add_argument('--target_bits', type=float, help='target bits', default=4)
    parser.add_argument('--arch_path', type=str, help='arch path', default=None)
    parser.add_argument('--file_name', type=str, help='save path', default=None)
    # additions
    parser.add_argument('--synthetic_arch', action='store_true',
                        help='no --arch_path: draw per-linear bits like the search space until bits_usage hits --target_bits')
    parser.add_argument('--n_block', type=int, default=None, help='build only the first n decoder blocks (quick runs)')
    parser.add_argument('--seed', type=int, default=0)
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    import torch
    from amq_b200.arch import MODELS, get_bits_usage, load_arch, sample_arch
    from amq_b200.model import QuantDecoder
    from amq_b200.utils.speed import benchmark_speed

    if args.model_name not in MODELS:
        raise KeyError(f"unknown model {args.model_name}; known shapes: {sorted(MODELS)}")
    if not torch.cuda.is_available():
        raise RuntimeError("amq_speed_benchmark: needs a CUDA device (no CPU path)")
    shape = MODELS[args.model_name]
    target_bits = args.target_bits
    if args.arch_path is None and args.synthetic_arch:
        arch = sample_arch(shape, target_bits, seed=args.seed)
    else:
        arch = load_arch(args.arch_path, target_bits, shape.n_block)
    result = {}
    sizes = [args.batch_size, args.seq_length, args.gen_length]
    gemm_iteration = 20
    gemv_iteration = 5 if args.gen_length < 1024 else 2

    print("Replacing...")
    model = QuantDecoder(shape, arch, batch=args.batch_size, max_seq=args.seq_length + args.gen_length + 8,
                         n_block=args.n_block, seed=args.seed)
    tag = f'{target_bits}bit'
    print(f"Get Speed of {target_bits}bit model...")
    result[tag] = {'bits_usage': get_bits_usage({'linear': arch}, shape.config())}
    if args.tps:
        tps = benchmark_speed(model, None, use_ft=args.use_ft, iteration=gemv_iteration, sizes=sizes, mode='TPS',
                              get_peak_memory=args.peak_memory)
        result[tag].update(tps)
        print('Token per second : ', tps)
    if args.gemm:
        gemm = benchmark_speed(model, None, use_ft=args.use_ft, iteration=gemm_iteration, sizes=sizes, mode='GeMM',
                               get_peak_memory=False)
        result[tag].update(gemm)
        print('GeMM : ', gemm)
    if args.gemv:
        gemv = benchmark_speed(model, None, use_ft=args.use_ft, iteration=gemv_iteration, sizes=sizes, mode='GeMV',
                               get_peak_memory=False)
        result[tag].update(gemv)
        print('GeMV : ', gemv)
    if args.ttft:
        ttft = benchmark_speed(model, None, use_ft=args.use_ft, iteration=gemm_iteration, sizes=sizes, mode='TTFT',
                               get_peak_memory=False)
        result[tag].update(ttft)
        print('TTFT : ', ttft)
    if args.memory:
        # get_memory_footprint (amq_speed_benchmark.py:90-97): parameters + buffers = packed linears + fp16 embed / head / norms
        memory = (model.weight_bytes + 2 * (model.embed.numel() + model.lm_head.numel())
                  + 2 * model.H * (2 * model.n_block + 1)) / 1024 ** 3
        result[tag].update({'memory': memory})
        print(f"Quantized Model Memory : {memory} GB")
    if args.file_name:
        result_dir = 'benchmark/outputs'
        os.makedirs(result_dir, exist_ok=True)
        with open(os.path.join(result_dir, args.file_name), 'w') as f:
            result.update({'args': vars(args)})
            json.dump(result, f, indent=4)
    return result


if __name__ == '__main__':
    main()
