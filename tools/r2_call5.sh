#!/bin/bash
mkdir -p gpurun_out
for v in r1prof prof_nosplit prof_split; do
  echo "=== $v" >> gpurun_out/c5_prof.txt
  S=75600 H=40 LITEATTN_B200_LIB=$PWD/tools/_build/lib_$v.so timeout 300 python tools/prof_clocks.py >> gpurun_out/c5_prof.txt 2>&1
done
cat gpurun_out/c5_prof.txt
