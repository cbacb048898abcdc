#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=1 ..."  ->  amq_b200/lib_NAME/libamqb.so (select at run time with AMQB_LIB=...)
name=$1; shift
AMQB_LIBDIR=amq_b200/lib_$name AMQB_CFLAGS="$*" python -m amq_b200.build --force
