#!/bin/bash
# Build a measurement variant of libfemb200.so with extra compiler defines for assemble.cu:
#   tools/build_variant.sh w16 -DFAN_WARPS4=16        ->  tools/micro/lib_w16.so
# Run it with FEMB200_LIB=tools/micro/lib_w16.so python bench.py ...   (fempy_b200/_lib.py honours FEMB200_LIB)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
python -m fempy_b200._build > /dev/null 2>&1
mkdir -p tools/micro
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c -o /tmp/assemble_$name.o fempy_b200/csrc/assemble.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/micro/lib_$name.so /tmp/assemble_$name.o fempy_b200/csrc/build/common.o fempy_b200/csrc/build/plan.o fempy_b200/csrc/build/functions.o fempy_b200/csrc/build/solve.o fempy_b200/csrc/build/generic.o
echo tools/micro/lib_$name.so
