#!/bin/bash
# usage (on the GPU box): tools/variants_run.sh "<lib names or 'default'>" "<bench arg sets ...>"
# runs tools/bench_rows.sh for each measurement build (tools/build_variant.sh) and argument set
libs=$1; shift
for l in $libs; do
  if [ "$l" = default ]; then unset FEMB200_LIB; else export FEMB200_LIB=tools/micro/lib_$l.so; fi
  tools/bench_rows.sh "$@"
done
