#!/bin/bash
# same-box comparison in the bench context: share of exp2 on the FMA pipe (2 pairs of 8 = old, none = p0, 1 = p1)
mkdir -p gpurun_out; : > gpurun_out/c32.txt
for rep in 1 2 3; do
  for v in old p0 p1; do
    LITEATTN_B200_LIB=$PWD/tools/_build/lib_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-comparators --no-traffic --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v rep $rep: step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'dense', round(d['sweep'][0]['fwd_ms'],3), 'bern', round(d['sweep'][1]['fwd_ms'],3), 's77', round(d['sweep'][2]['fwd_ms'],3), 'clk', d['clocks']['sm_mhz'])" >> gpurun_out/c32.txt
  done
done
cat gpurun_out/c32.txt
