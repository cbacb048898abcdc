#!/bin/bash
# One GPU-box visit: parity tests, bench, per-launch-class breakdown.  Every step has its own timeout and streams its
# output to gpurun_out/ (a hung step must not take the visit down).
tag=${1:-x}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tee gpurun_out/r02_pytest_$tag.txt | tail -${2:-12}
AMQB_SKIP_TP70B=1 timeout 300 python bench.py --steps 64 --warmup 8 > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err
timeout 200 python tools/model_breakdown.py > gpurun_out/r02_breakdown_$tag.txt 2>&1
AMQB_NO_PAIR=1 timeout 200 python tools/model_breakdown.py > gpurun_out/r02_breakdown_nopair_$tag.txt 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_$tag.json"))
print("bench:", d["value"], "tok/s  e2e", d["e2e"]["value"], " roofline", d["roofline"]["frac"], d["roofline"]["avg_launch_us"])
PY
tail -3 gpurun_out/r02_bench_$tag.err
cat gpurun_out/r02_breakdown_$tag.txt; echo "--- nopair"; cat gpurun_out/r02_breakdown_nopair_$tag.txt
