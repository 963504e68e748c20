#!/bin/bash
# usage: tools/bench_rows.sh "<bench args>" ... ; prints step / kernel time for each argument set
for a in "$@"; do
  timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu $a 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$a', 'step_us', round(1e3*d['ms_per_step'],1), 'kernel_us', round(1e3*d['roofline']['kernel_ms'],1), 'frac', round(d['roofline']['frac'],3))"
done
