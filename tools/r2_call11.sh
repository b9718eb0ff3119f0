#!/bin/bash
mkdir -p gpurun_out
LITEATTN_B200_LIB=$PWD/tools/_build/lib_r1.so ncu --set full --clock-control none --import-source on -k regex:la_fwd_kernel -s 2 -c 1 -o gpurun_out/prof_cmp_r1 -f python tools/one_launch.py > gpurun_out/c11_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:la_fwd_kernel -s 2 -c 1 -o gpurun_out/prof_cmp_pers -f python tools/one_launch.py > gpurun_out/c11_b.log 2>&1
ls -la gpurun_out/prof_cmp_*; tail -3 gpurun_out/c11_b.log
