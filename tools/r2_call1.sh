#!/bin/bash
# round-2 GPU call 1: parity suite, softmax organisation microbenchmark, dense comparators, same-box baseline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 300 tools/_build/softmax_rate2 3000 > gpurun_out/softmax_rate2.txt 2>&1
timeout 600 python - > gpurun_out/comparators.json 2> gpurun_out/comparators.err <<'PY'
import json, sys, torch
sys.path.insert(0, ".")
from baseline.comparators import time_dense_comparators
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(1, 75600, 40, 128, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
print(json.dumps(time_dense_comparators(q, k, v, steps=5, warmup=2), indent=1))
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --sweep > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -3 gpurun_out/c1_pytest.log; cat gpurun_out/softmax_rate2.txt; cat gpurun_out/comparators.json
