#!/bin/bash
# l correction for the largest P (LA_LMAX_FIX) and a larger lazy bound: accuracy on peaked rows, parity tests, C3 and bench timing
mkdir -p gpurun_out; : > gpurun_out/c40.txt
for v in base tau64 fix8 fix64; do
  echo "== $v" >> gpurun_out/c40.txt
  LITEATTN_B200_LIB=$PWD/tools/_build/lib_$v.so timeout 300 python tools/acc_probe.py >> gpurun_out/c40.txt 2>&1
done
for v in fix8 fix64; do
  LITEATTN_B200_LIB=$PWD/tools/_build/lib_$v.so timeout 900 python -m pytest tests/test_fwd_gpu.py tests/test_combine_gpu.py -m gpu -x -q > gpurun_out/c40_pytest_$v.log 2>&1; echo "$v pytest: $(tail -1 gpurun_out/c40_pytest_$v.log)" >> gpurun_out/c40.txt
done
for v in base fix8 fix64 base fix64; do
  LITEATTN_B200_LIB=$PWD/tools/_build/lib_$v.so timeout 600 python tools/c3_trajectory.py 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v C3: mean', d['mean_ms'], 'ms at step', d['ms_at_step'], 'final sparsity', d['final_list_sparsity'])" >> gpurun_out/c40.txt
done
for rep in 1 2; do
  for v in base fix8 fix64; do
    LITEATTN_B200_LIB=$PWD/tools/_build/lib_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-comparators --no-traffic --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v rep $rep: step', round(d['ms_per_step'],3), 'dense', round(d['sweep'][0]['fwd_ms'],3), 'bern', round(d['sweep'][1]['fwd_ms'],3), 's77', round(d['sweep'][2]['fwd_ms'],3), 'clk', d['clocks']['sm_mhz'])" >> gpurun_out/c40.txt
  done
done
cat gpurun_out/c40.txt
