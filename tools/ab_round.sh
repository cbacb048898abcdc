#!/bin/bash
# tools/ab_round.sh "ENV=.. ENV=.." lib  (one line per configuration on stdin: "<label> <libname|default> [ENV=VAL ...]")
while read -r label libn envs; do
  [ -z "$label" ] && continue
  if [ "$libn" = default ]; then lib=""; else lib=$PWD/amq_b200/lib_$libn/libamqb.so; fi
  echo "=== $label ($libn $envs)"
  env AMQB_LIB=$lib $envs timeout 200 python tools/model_breakdown.py 2>&1 | tail -7 | tr '\n' ';' | sed 's/ us per launch//g; s/  */ /g'
  echo
done
