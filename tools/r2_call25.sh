#!/bin/bash
# host-streamed calls (tests + bench e2e), L2 policy variants (DRAM traffic + time), ncu of the bitmap update kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_stream_gpu.py tests/test_update_gpu.py -m gpu -x -q > gpurun_out/c25_pytest.log 2>&1; tail -5 gpurun_out/c25_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-comparators --no-traffic > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c25_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c25_bench.json').read().strip().splitlines()[-1])
print("ms/step",d['ms_per_step'],"e2e",d['e2e']['ms_per_step'], d['e2e']['api'])
print(d['e2e']['host_link']); print(d['roofline']['update_kernel'])
PY
for v in h0 h2 h4 h6 h3 h7; do
  echo "== $v" >> gpurun_out/c25_traffic.txt
  LITEATTN_B200_LIB=$PWD/tools/_build/lib_$v.so timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:la_fwd --launch-skip 2 -c 1 python tools/one_launch.py 2>&1 | grep -E 'dram__|gpu__time' >> gpurun_out/c25_traffic.txt
done
cat gpurun_out/c25_traffic.txt
timeout 1200 python tools/ab.py --rounds 1 --secs 1.5 h0=tools/_build/lib_h0.so h2=tools/_build/lib_h2.so h4=tools/_build/lib_h4.so h6=tools/_build/lib_h6.so h3=tools/_build/lib_h3.so h7=tools/_build/lib_h7.so > gpurun_out/c25_ab.txt 2>&1
tail -8 gpurun_out/c25_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:la_skip_update -c 1 -o gpurun_out/prof_upd_bitmap_r2 -f python tools/one_launch.py --update > gpurun_out/prof_upd_bitmap_r2.log 2>&1; tail -2 gpurun_out/prof_upd_bitmap_r2.log
