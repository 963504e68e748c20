#!/bin/bash
# usage: tools/bench_rows.sh "<bench args>" ... ; prints step / kernel time for each argument set
# (FEMB200_LIB=tools/micro/lib_<variant>.so selects a measurement build, tools/build_variant.sh)
for a in "$@"; do
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu $a 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('${FEMB200_LIB##*/} $a', 'step_us', round(1e3*d['ms_per_step'],1), 'min_us', round(1e3*d['ms_per_step_min'],1), 'kernel_us', round(1e3*d['roofline']['kernel_ms'],1), 'frac', round(d['roofline']['frac'],3), 'fan', d['path'].get('fan_path'))"
done
