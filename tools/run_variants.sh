#!/bin/bash
# usage: tools/run_variants.sh k0 k6 ... : bench each tools/micro/lib_<v>.so (kernel time on config 2)
for v in "$@"; do cp tools/micro/lib_$v.so fempy_b200/csrc/libfemb200.so; echo -n "$v: "; tools/bench_rows.sh "--path rows" | sed 's/--path rows//'; done
