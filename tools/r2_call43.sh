#!/bin/bash
# mbarrier.try_wait with a suspend-time hint (fewer polls by the issuer / producer / waiting softmax warps): bench-context A/B
mkdir -p gpurun_out; : > gpurun_out/c43.txt
LITEATTN_B200_LIB=$PWD/tools/_build/lib_hint1k.so timeout 600 python -m pytest tests/test_fwd_gpu.py -m gpu -x -q 2>&1 | tail -1 >> gpurun_out/c43.txt
for rep in 1 2 3; do
  for v in base hint1k hint20k; do
    LITEATTN_B200_LIB=$PWD/tools/_build/lib_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-comparators --no-traffic --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v rep $rep: step', round(d['ms_per_step'],3), 'dense', round(d['sweep'][0]['fwd_ms'],3), 'bern', round(d['sweep'][1]['fwd_ms'],3), 's77', round(d['sweep'][2]['fwd_ms'],3), 'clk', d['clocks']['sm_mhz'], 'W', d['clocks'].get('power_w_max'))" >> gpurun_out/c43.txt
  done
done
cat gpurun_out/c43.txt
