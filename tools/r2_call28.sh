#!/bin/bash
# half-row max exchange through tagged smem slots (no named barrier in the hot loop): correctness on one variant, A/B
mkdir -p gpurun_out
LITEATTN_B200_LIB=$PWD/tools/_build/lib_x18.so timeout 900 python -m pytest tests/test_fwd_gpu.py tests/test_ref_softmax_gpu.py -m gpu -x -q > gpurun_out/c28_pytest.log 2>&1; tail -3 gpurun_out/c28_pytest.log
timeout 1500 python tools/ab.py --rounds 2 --secs 1.5 base=tools/_build/lib_base.so x22=tools/_build/lib_x22.so x18=tools/_build/lib_x18.so x14=tools/_build/lib_x14.so x10=tools/_build/lib_x10.so > gpurun_out/c28_ab.txt 2>&1
tail -7 gpurun_out/c28_ab.txt
