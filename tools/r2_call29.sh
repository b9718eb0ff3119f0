#!/bin/bash
# tagged exchange + verdict attempted inside the loop with a predicated early P store: correctness + A/B
mkdir -p gpurun_out
LITEATTN_B200_LIB=$PWD/tools/_build/lib_e14_18.so timeout 900 python -m pytest tests/test_fwd_gpu.py tests/test_ref_softmax_gpu.py -m gpu -x -q > gpurun_out/c29_pytest.log 2>&1; tail -3 gpurun_out/c29_pytest.log
timeout 1500 python tools/ab.py --rounds 2 --secs 1.5 base=tools/_build/lib_base.so x18=tools/_build/lib_x18.so e14_18=tools/_build/lib_e14_18.so e14_20=tools/_build/lib_e14_20.so e10_16=tools/_build/lib_e10_16.so e16_20=tools/_build/lib_e16_20.so > gpurun_out/c29_ab.txt 2>&1
tail -8 gpurun_out/c29_ab.txt
