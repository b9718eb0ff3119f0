#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fwd_gpu.py -m gpu -x -q -k "dense or list_gated or multi or wan" > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c12_pytest.log
tail -3 gpurun_out/c12_pytest.log
timeout 900 python tools/ab.py --rounds 1 --secs 1.0 r1=tools/_build/lib_r1.so r1tw=tools/_build/lib_r1tw.so pers=- pers_trywait=tools/_build/lib_trywait.so > gpurun_out/c12_ab.txt 2>&1
cat gpurun_out/c12_ab.txt
S=75600 H=40 LITEATTN_B200_LIB=$PWD/tools/_build/lib_prof.so timeout 300 python tools/prof_clocks.py > gpurun_out/c12_prof.txt 2>&1
cat gpurun_out/c12_prof.txt
