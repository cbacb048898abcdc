#!/bin/bash
# Round-2 measurement set in one GPU-box call (every step with its own timeout, outputs under gpurun_out/r02f_*):
# parity, bench (both arms), per-class breakdown, ncu launch list + full capture of one layer's GEMV launches, kernel-level
# microbench, configs 3 / 4, the reference's protocol, the module path.
o=gpurun_out; mkdir -p $o
timeout 700 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4 > $o/r02f_pytest_gpu.txt; cat $o/r02f_pytest_gpu.txt
timeout 600 python bench.py 2> $o/r02f_bench.err | tail -1 > $o/r02f_bench.json; cut -c1-400 $o/r02f_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $o/r02f_bench_reference_arm.json; cut -c1-200 $o/r02f_bench_reference_arm.json
timeout 200 python tools/model_breakdown.py > $o/r02f_model_breakdown.txt 2>&1; cat $o/r02f_model_breakdown.txt
AMQB_SKIP_TP70B=1 AMQB_PROFILE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r02f_launches_bench.csv python bench.py --steps 2 --warmup 3 > $o/r02f_ncu_bench.log 2>&1
python tools/launch_summary.py $o/r02f_launches_bench.csv > $o/r02f_launches_bench_summary.txt 2>&1; head -14 $o/r02f_launches_bench_summary.txt
AMQB_SKIP_TP70B=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:gemv_mma -s 40 -c 4 -o $o/r02f_gemv_layer -f python bench.py --steps 2 --warmup 3 > $o/r02f_ncu_full.log 2>&1
python tools/ncu_traffic.py $o/r02f_gemv_layer.ncu-rep $o/r02f_ncu_gemv_traffic.json > /dev/null 2>&1; head -c 600 $o/r02f_ncu_gemv_traffic.json; echo
timeout 300 python tools/microbench.py --out $o/r02f_microbench_gemv.jsonl > $o/r02f_microbench.log 2>&1; grep '"pdl": true' $o/r02f_microbench.log | grep '"M": 1,' | cut -c1-110
timeout 500 python tools/run_configs.py > $o/r02f_run_configs.log 2>&1; cp $o/configs.json $o/r02f_configs_3_4.json 2>/dev/null; tail -6 $o/r02f_run_configs.log | cut -c1-300
timeout 300 python tools/ref_protocol.py --iters 3 --out $o/r02f_ref_protocol.json 2>&1 | tail -1 | cut -c1-400
timeout 300 python amq_speed_benchmark.py --target_bits 3 --synthetic_arch --tps --gemv --gemm --ttft --memory --peak_memory > $o/r02f_speed_benchmark_7b.txt 2>&1; tail -8 $o/r02f_speed_benchmark_7b.txt
