#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fwd_gpu.py -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
tail -3 gpurun_out/c4_pytest.log
timeout 1200 python tools/ab.py --rounds 2 --secs 1.0 r1=tools/_build/lib_r1.so new=- tc=tools/_build/lib_tc.so nohint=tools/_build/lib_nohint.so nosplit=tools/_build/lib_nosplit.so nosplit_nohint=tools/_build/lib_nosplit_nohint.so > gpurun_out/c4_ab.txt 2>&1
cat gpurun_out/c4_ab.txt
