#!/bin/bash
# 8 GPUs: bench N=8 (weak scaling, fused gather, sequence-parallel leg, e2e through the host-resident call)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/c34_bench_n8.json 2> gpurun_out/c34_bench_n8.err; echo "bench rc=$?"
tail -3 gpurun_out/c34_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c34_bench_n8.json').read().strip().splitlines()[-1])
print("N=8 ms/step",d['ms_per_step'],"value",d['value'],"gather_verified",d.get('gather_verified'))
print("seq_parallel",d.get('seq_parallel'))
print("e2e",d.get('e2e'))
PY
