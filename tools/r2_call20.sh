#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_update_gpu.py -m gpu -x -q > gpurun_out/c20_pytest.log 2>&1; tail -2 gpurun_out/c20_pytest.log
echo "== new" > gpurun_out/c20_upd.txt; timeout 300 python tools/time_update.py >> gpurun_out/c20_upd.txt 2>&1
echo "== r1" >> gpurun_out/c20_upd.txt; LITEATTN_B200_LIB=$PWD/tools/_build/lib_r1.so timeout 300 python tools/time_update.py >> gpurun_out/c20_upd.txt 2>&1
cat gpurun_out/c20_upd.txt
