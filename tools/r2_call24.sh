#!/bin/bash
# bitmap-domain update kernel + host-streamed calls: parity tests, update timing vs the previous build, bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_update_gpu.py tests/test_compact_state_gpu.py tests/test_host_stream_gpu.py tests/test_fwd_gpu.py -m gpu -x -q > gpurun_out/c24_pytest.log 2>&1; tail -15 gpurun_out/c24_pytest.log
echo "== new" > gpurun_out/c24_upd.txt; timeout 300 python tools/time_update.py >> gpurun_out/c24_upd.txt 2>&1
echo "== previous" >> gpurun_out/c24_upd.txt; LITEATTN_B200_LIB=$PWD/tools/_build/lib_updold.so timeout 300 python tools/time_update.py >> gpurun_out/c24_upd.txt 2>&1
cat gpurun_out/c24_upd.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-comparators --no-traffic > gpurun_out/c24_bench.json 2> gpurun_out/c24_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c24_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c24_bench.json').read().strip().splitlines()[-1])
print("ms/step",d['ms_per_step'],"e2e",d['e2e']['ms_per_step'], d['e2e']['api'])
print(d['e2e']['host_link']); print(d['roofline']['update_kernel'])
PY
for H in 1 2 40; do
  for rep in 1 2; do
    echo "== heads $H rep $rep" >> gpurun_out/c24_traffic.txt
    timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_ltcfabric.sum --clock-control none -k regex:la_fwd --launch-skip 2 -c 1 python tools/one_launch.py --heads $H 2>&1 | grep -E 'dram__|gpu__time|ltcfabric' >> gpurun_out/c24_traffic.txt
  done
done
cat gpurun_out/c24_traffic.txt
