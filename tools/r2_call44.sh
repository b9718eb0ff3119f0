#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_stream_gpu.py -m gpu -x -q > gpurun_out/c44_pytest.log 2>&1; tail -5 gpurun_out/c44_pytest.log
