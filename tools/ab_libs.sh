#!/bin/bash
# tools/ab_libs.sh name1 name2 ...: per-launch-class breakdown of the 7B decode step for each in-tree library variant
# ("default" = amq_b200/lib).  Extra environment (AMQB_*) is passed through.
for v in "$@"; do
  if [ "$v" = default ]; then lib=""; else lib=$PWD/amq_b200/lib_$v/libamqb.so; fi
  echo "=== $v"
  AMQB_LIB=$lib timeout 200 python tools/model_breakdown.py 2>&1 | tail -8
done
