#!/bin/bash
# same-box comparison in the bench context (forward + update through the public objects): named barrier (old) vs tagged slots (new)
mkdir -p gpurun_out; : > gpurun_out/c31.txt
for rep in 1 2 3; do
  for v in old new x14; do
    LITEATTN_B200_LIB=$PWD/tools/_build/lib_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-comparators --no-traffic --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v rep $rep: step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'dense', round(d['sweep'][0]['fwd_ms'],3), 'bern', round(d['sweep'][1]['fwd_ms'],3), 's77', round(d['sweep'][2]['fwd_ms'],3), 'clk', d['clocks']['sm_mhz'])" >> gpurun_out/c31.txt
  done
done
cat gpurun_out/c31.txt
