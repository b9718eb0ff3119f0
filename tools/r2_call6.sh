#!/bin/bash
mkdir -p gpurun_out
timeout 200 tools/_build/softmax_rate2 1500 2>&1 | head -8 > gpurun_out/softmax_rate2c.txt
cat gpurun_out/softmax_rate2c.txt
