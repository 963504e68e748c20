#!/bin/bash
# Build measurement variants of the row-owner kernel with stages stubbed out (ROWS_KNOCK bit mask, assemble_rows.cuh)
# into tools/micro/lib_k<mask>.so.  usage: tools/knock_variants.sh 0 6 41 ...
set -e
cd "$(dirname "$0")/.."
python fempy_b200/_build.py > /dev/null
mkdir -p tools/micro
for k in "$@"; do
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DROWS_KNOCK=$k -c -o /tmp/assemble_k$k.o fempy_b200/csrc/assemble.cu &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/micro/lib_k$k.so /tmp/assemble_k$k.o fempy_b200/csrc/build/common.o fempy_b200/csrc/build/plan.o fempy_b200/csrc/build/plan_patch.o fempy_b200/csrc/build/functions.o ) &
done
wait
ls -la tools/micro/*.so
