#!/bin/bash
mkdir -p gpurun_out
timeout 240 tools/_build/softmax_rate2 2000 > gpurun_out/softmax_rate2b.txt 2>&1; echo "rc=$?" >> gpurun_out/softmax_rate2b.txt
timeout 600 python -m pytest tests/test_ref_softmax_gpu.py tests/test_combine_gpu.py tests/test_fwd_gpu.py -m gpu -q -k "ref or combine or wan_shape or multi_timestep or list_gated" > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
cat gpurun_out/softmax_rate2b.txt; tail -15 gpurun_out/c2_pytest.log
