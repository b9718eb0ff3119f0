#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab.py --rounds 1 --secs 1.0 r1=tools/_build/lib_r1.so pers=- nov8=tools/_build/lib_nov8.so nohint=tools/_build/lib_nohint.so nov8nohint=tools/_build/lib_nov8nohint.so > gpurun_out/c13_ab.txt 2>&1
cat gpurun_out/c13_ab.txt
for v in r1 nov8 nohint nov8nohint; do
  echo "== $v" >> gpurun_out/c13_traffic.txt
  LITEATTN_B200_LIB=$PWD/tools/_build/lib_$v.so ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:la_fwd_kernel -s 2 -c 1 python tools/one_launch.py 2>&1 | grep -E "dram__|duration" >> gpurun_out/c13_traffic.txt
done
echo "== pers" >> gpurun_out/c13_traffic.txt
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:la_fwd_kernel -s 2 -c 1 python tools/one_launch.py 2>&1 | grep -E "dram__|duration" >> gpurun_out/c13_traffic.txt
cat gpurun_out/c13_traffic.txt
