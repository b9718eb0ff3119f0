#!/bin/bash
# Run under gpurun (1 GPU): ncu launch list of the bench command + one --set full capture per kernel.
#   tools/capture_profiles.sh <tag>        -> gpurun_out/launches_<tag>.csv, prof_{fwd,upd,cmb}_<tag>.ncu-rep
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-comparators --no-traffic"
# every launch of the bench command with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:la_ -c 12 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/launches_$TAG.log 2>&1
# forward kernel at the Wan2.1-14B shape, 42 % sparsity (config C3/C4 direct variant)
ncu --set full --clock-control none --import-source on -k regex:la_fwd_kernel -s 3 -c 1 -o $OUT/prof_fwd_wan42_$TAG -f $BENCH > $OUT/prof_fwd_wan42_$TAG.log 2>&1
# forward kernel, config C2 (S=32768, H=16, fixed random 50 % mask)
ncu --set full --clock-control none --import-source on -k regex:la_fwd_kernel -s 3 -c 1 -o $OUT/prof_fwd_c2_$TAG -f python bench.py --seq 32768 --heads 16 --sparsity 0.5 --steps 2 --warmup 3 --no-e2e --no-cpu --no-comparators --no-traffic > $OUT/prof_fwd_c2_$TAG.log 2>&1
# skip-list update kernel at the Wan shape
ncu --set full --clock-control none --import-source on -k regex:la_skip_update_kernel -s 3 -c 1 -o $OUT/prof_upd_wan42_$TAG -f $BENCH > $OUT/prof_upd_wan42_$TAG.log 2>&1
# fused RoPE + cast kernel (runs inside the bench's aux timing)
ncu --set full --clock-control none --import-source on -k regex:la_rope_cast -s 3 -c 1 -o $OUT/prof_rope_wan_$TAG -f $BENCH > $OUT/prof_rope_wan_$TAG.log 2>&1
ls -la $OUT | tail -12
