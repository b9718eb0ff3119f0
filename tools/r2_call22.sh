#!/bin/bash
mkdir -p gpurun_out
bash tools/capture_profiles.sh r2 > gpurun_out/c22_capture.log 2>&1
tail -8 gpurun_out/c22_capture.log
# dense launch too (the metric's 0 % point)
ncu --set full --clock-control none --import-source on -k regex:la_fwd_kernel -s 2 -c 1 -o gpurun_out/prof_fwd_wan00_r2 -f python tools/one_launch.py --sparsity 0 > gpurun_out/c22_dense.log 2>&1
# update kernel with a must-do list
timeout 300 python tools/time_update.py > gpurun_out/update_timing_r2.txt 2>&1; cat gpurun_out/update_timing_r2.txt
# cycle accounting of the shipped kernel
python tools/build_variants.py prof=LA_PROFILE_CLOCKS > /dev/null 2>&1
S=75600 H=40 LITEATTN_B200_LIB=$PWD/tools/_build/lib_prof.so timeout 300 python tools/prof_clocks.py > gpurun_out/prof_clocks_r2.txt 2>&1; cat gpurun_out/prof_clocks_r2.txt
