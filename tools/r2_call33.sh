#!/bin/bash
# 2 GPUs, shipped library: whole GPU suite (incl. the two multi-GPU tests) + bench N=2
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/c33_pytest.log 2>&1; tail -3 gpurun_out/c33_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c33_bench_n2.json 2> gpurun_out/c33_bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/c33_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c33_bench_n2.json').read().strip().splitlines()[-1])
print("N=2 ms/step",d['ms_per_step'],"value",d['value'],"gather_verified",d.get('gather_verified'))
print("seq_parallel",d.get('seq_parallel'))
print("e2e",d.get('e2e'))
PY
