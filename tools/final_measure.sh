#!/bin/bash
# Round measurements in one GPU-box call: parity, bench (both arms), ncu launch list + full capture, microbench,
# per-launch-class breakdown.  Everything lands in gpurun_out/ (copied to profiles/ afterwards).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python bench.py 2>&1 | tail -1 > gpurun_out/bench.json; cut -c1-300 gpurun_out/bench.json
python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/bench_reference_arm.json; cut -c1-300 gpurun_out/bench_reference_arm.json
python tools/model_breakdown.py > gpurun_out/model_breakdown.txt 2>&1; cat gpurun_out/model_breakdown.txt
AMQB_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bench.csv > gpurun_out/launches_bench_summary.txt 2>&1; head -12 gpurun_out/launches_bench_summary.txt
ncu --set full --clock-control none --import-source on -k regex:gemv_mma -s 40 -c 4 -o gpurun_out/gemv_layer -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
python tools/ncu_traffic.py gpurun_out/gemv_layer.ncu-rep gpurun_out/ncu_gemv_traffic.json > /dev/null 2>&1; head -c 600 gpurun_out/ncu_gemv_traffic.json
python tools/microbench.py --out gpurun_out/microbench_gemv.jsonl > gpurun_out/microbench.log 2>&1; grep '"pdl": true' gpurun_out/microbench.log | grep '"M": 1,' | cut -c1-100
