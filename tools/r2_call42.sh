#!/bin/bash
# maximum sizes: 2048 K tiles through the forward + update, update kernel with more than one 32-word lane group
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_update_gpu.py tests/test_fwd_gpu.py -m gpu -x -q > gpurun_out/c42_pytest.log 2>&1; tail -15 gpurun_out/c42_pytest.log
