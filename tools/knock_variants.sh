#!/bin/bash
# Build measurement variants of the row-owner kernel with stages stubbed out (ROWS_KNOCK bit mask, assemble_rows.cuh)
# into tools/micro/lib_k<mask>.so; "trace" builds lib_trace.so with warp-lifetime statistics (tools/rows_trace.py).
# usage: tools/knock_variants.sh 0 6 41 v1 trace ...   then, on the GPU box: tools/run_variants.sh k0 k6 v1 ...
set -e
cd "$(dirname "$0")/.."
python fempy_b200/_build.py > /dev/null
mkdir -p tools/micro
for k in "$@"; do
  if [ "$k" = trace ]; then def="-DROWS_TRACE=1"; out=lib_trace.so;
  elif [ "${k#v}" != "$k" ]; then def="-DROWS_VAR=${k#v}"; out=lib_$k.so;   # v<mask>: spare switch for experimental schedules (define ROWS_VAR tests in the kernel)
  else def="-DROWS_KNOCK=$k"; out=lib_k$k.so; fi
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $def -c -o /tmp/assemble_k$k.o fempy_b200/csrc/assemble.cu &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/micro/$out /tmp/assemble_k$k.o fempy_b200/csrc/build/common.o fempy_b200/csrc/build/plan.o fempy_b200/csrc/build/plan_patch.o fempy_b200/csrc/build/functions.o ) &
done
wait
ls -la tools/micro/*.so
