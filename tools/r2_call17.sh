#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab.py --rounds 1 --secs 1.0 r1=tools/_build/lib_r1.so pers=- > gpurun_out/c17_ab.txt 2>&1
LA_FWD_GRID=1000000 timeout 900 python tools/ab.py --rounds 1 --secs 1.0 pers_oneshot=- >> gpurun_out/c17_ab.txt 2>&1
LA_FWD_GRID=296 timeout 900 python tools/ab.py --rounds 1 --secs 1.0 pers_296=- >> gpurun_out/c17_ab.txt 2>&1
cat gpurun_out/c17_ab.txt
