#!/bin/bash
# early-verdict / progressive P store variants: correctness of one variant, then same-box A/B
mkdir -p gpurun_out
LITEATTN_B200_LIB=$PWD/tools/_build/lib_e2p8.so timeout 900 python -m pytest tests/test_fwd_gpu.py -m gpu -x -q > gpurun_out/c23_pytest.log 2>&1; tail -3 gpurun_out/c23_pytest.log
timeout 1500 python tools/ab.py --rounds 2 --secs 1.5 base=tools/_build/lib_base.so e1p8=tools/_build/lib_e1p8.so e1p12=tools/_build/lib_e1p12.so e2p8=tools/_build/lib_e2p8.so e2p4=tools/_build/lib_e2p4.so e1p15=tools/_build/lib_e1p15.so > gpurun_out/c23_ab.txt 2>&1
cat gpurun_out/c23_ab.txt
