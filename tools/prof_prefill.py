"""One eager 63-row prompt pass of Llama-2-7B @ 3.0 bits (4 blocks) between cudaProfilerStart / Stop, for an ncu launch
list: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/prof_prefill.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from amq_b200.arch import MODELS, sample_arch
from amq_b200.model import QuantDecoder

shape = MODELS["Llama-2-7b-hf"]
arch = sample_arch(shape, 3.0, seed=0)
m = QuantDecoder(shape, arch, batch=1, max_seq=128, n_block=int(os.environ.get("AMQB_BLOCKS", "4")))
ids = torch.randint(0, shape.vocab, (1, int(os.environ.get("ROWS", "63"))), device=m.dev)
for _ in range(2):
    m.reset()
    m.prefill(ids, use_graph=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m.reset()
m.prefill(ids, use_graph=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
