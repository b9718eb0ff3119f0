#!/bin/bash
# Q + first K tile requested before the list decode: parity tests, then same-box comparison inside bench.py
mkdir -p gpurun_out; : > gpurun_out/c36.txt
timeout 1200 python -m pytest tests/test_fwd_gpu.py tests/test_host_stream_gpu.py tests/test_dist_gpu.py -m gpu -x -q > gpurun_out/c36_pytest.log 2>&1; tail -3 gpurun_out/c36_pytest.log
for rep in 1 2 3; do
  for v in noearly early; do
    LITEATTN_B200_LIB=$PWD/tools/_build/lib_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-comparators --no-traffic --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v rep $rep: step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'dense', round(d['sweep'][0]['fwd_ms'],3), 'bern', round(d['sweep'][1]['fwd_ms'],3), 's77', round(d['sweep'][2]['fwd_ms'],3), 'clk', d['clocks']['sm_mhz'])" >> gpurun_out/c36.txt
  done
done
cat gpurun_out/c36.txt
