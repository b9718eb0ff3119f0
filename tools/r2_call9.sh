#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/dbg_small.py > gpurun_out/c9_dbg.txt 2>&1; echo "dbg rc=$?" >> gpurun_out/c9_dbg.txt; tail -3 gpurun_out/c9_dbg.txt
S=5000 H=4 timeout 120 python tools/dbg_small.py >> gpurun_out/c9_dbg.txt 2>&1; tail -2 gpurun_out/c9_dbg.txt
timeout 900 python -m pytest tests/test_fwd_gpu.py tests/test_ref_softmax_gpu.py tests/test_update_gpu.py -m gpu -x -q > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
tail -12 gpurun_out/c9_pytest.log
timeout 600 python tools/ab.py --rounds 1 --secs 1.0 r1=tools/_build/lib_r1.so pers=- > gpurun_out/c9_ab.txt 2>&1
cat gpurun_out/c9_ab.txt
S=75600 H=40 LITEATTN_B200_LIB=$PWD/tools/_build/lib_prof.so timeout 300 python tools/prof_clocks.py > gpurun_out/c9_prof.txt 2>&1
cat gpurun_out/c9_prof.txt
