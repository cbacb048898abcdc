#!/bin/bash
# Opcode histogram of the shipped library (what proves the Blackwell-native paths: UTCHMMA = tcgen05.mma,
# LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA engine), IMMA = mma.sync u8 x s8, SYNCS = mbarrier).
#   tools/sass_opcodes.sh > profiles/r02_sass_opcodes.txt
LIB=${1:-amq_b200/lib/libamqb.so}
echo "# cuobjdump -sass $LIB | opcode histogram ($(date -u +%Y-%m-%d), $(nvcc --version | tail -1))"
echo "## whole library"
cuobjdump -sass "$LIB" | grep -oE '^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T]+ )?[A-Z][A-Z0-9_]*(\.[A-Z0-9_]+)*' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+ )?//' | sort | uniq -c | sort -rn | awk '$2 ~ /^(UTC|LDTM|STTM|UBLKCP|UTMA|IMMA|HMMA|SYNCS|REDUX|CREDUX|LDGSTS|UCGABAR|MEMBAR|ATOM|RED|LDG|STG|LDS|STS|MUFU|BAR|ELECT|NANOSLEEP|ACQBULK|FENCE|LOP3|PRMT|FFMA|HFMA2|IADD3|SHFL)/'
echo "## per kernel: tensor / TMA / mbarrier opcodes"
for k in $(cuobjdump -sass "$LIB" | grep -oE 'Function : [A-Za-z0-9_]+' | awk '{print $3}'); do
  n=$(cuobjdump -sass -fun "$k" "$LIB" 2>/dev/null | grep -oE '(UTC[A-Z]*MMA|LDTM|STTM|UBLKCP|IMMA\.[0-9]+|HMMA\.[0-9]+|UTCBAR|SYNCS\.[A-Z]+)' | sort | uniq -c | awk '{printf "%s=%s ", $2, $1}')
  [ -n "$n" ] && echo "$(echo $k | c++filt | cut -c1-110): $n"
done
