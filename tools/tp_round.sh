#!/bin/bash
# tools/tp_round.sh N: tensor-parallel parity test + config 5 (70B) on N GPUs with the three all-reduce kinds
N=${1:-2}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_tp.py -m gpu -q --timeout 400 -x 2>&1 | tail -5
for kind in fused amqb nccl; do
  AMQB_AR=$kind timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --workload llama70b-tp --gpus $N --steps 48 --warmup 6 2>gpurun_out/tp_${N}_$kind.err | tail -1 | tee -a gpurun_out/r02_tp70b_$N.jsonl | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$kind', 'tp', d['n_gpus'], 'tok/s %.1f' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'launches', d['config']['launches_per_step'], 'frac %.3f' % d['config']['frac_of_hbm_roofline_per_gpu'])"
  tail -2 gpurun_out/tp_${N}_$kind.err | cut -c1-300
done
