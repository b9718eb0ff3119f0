#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
tail -5 gpurun_out/c3_pytest.log
timeout 900 python tools/ab.py --rounds 2 --secs 1.5 r1=tools/_build/lib_r1.so new=- > gpurun_out/c3_ab.txt 2>&1
cat gpurun_out/c3_ab.txt
