#!/bin/bash
# share of the exponentials on the FMA pipe (polynomial pairs of every 8): 0, 1, 1.5, 2 (shipped), 3
mkdir -p gpurun_out
timeout 1200 python tools/ab.py --rounds 2 --secs 1.5 base=tools/_build/lib_base.so p0=tools/_build/lib_p0.so p1=tools/_build/lib_p1.so p15=tools/_build/lib_p15.so p3=tools/_build/lib_p3.so > gpurun_out/c27_ab.txt 2>&1
tail -7 gpurun_out/c27_ab.txt
