#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c18_pytest.log
tail -5 gpurun_out/c18_pytest.log
timeout 900 python tools/ab.py --rounds 2 --secs 1.0 r1=tools/_build/lib_r1.so new=- nohint=tools/_build/lib_nohint.so nov8=tools/_build/lib_nov8.so nohint_nov8=tools/_build/lib_nohint_nov8.so > gpurun_out/c18_ab.txt 2>&1
cat gpurun_out/c18_ab.txt
