#!/bin/bash
# 8 GPUs, host-link-bound e2e: head-group schedules of the host-resident call (default: coarse while busy; fine always; uniform 8 x 5)
mkdir -p gpurun_out; : > gpurun_out/c38.txt
for sched in default 2,4,8,13,9,4 5,5,5,5,5,5,5,5 1,2,4,8,10,10,4,1; do
  if [ "$sched" = default ]; then unset LITE_ATTENTION_HOST_CHUNKS; else export LITE_ATTENTION_HOST_CHUNKS=$sched; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 8 --steps 10 --warmup 3 --no-seqpar --no-cpu 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$sched: step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'floor', round(d['e2e']['host_link']['copy_only_floor_ms_per_step'],1), 'h2d/d2h', round(d['e2e']['host_link']['aggregate_h2d_gbs']), round(d['e2e']['host_link']['aggregate_d2h_gbs']))" >> gpurun_out/c38.txt
done
cat gpurun_out/c38.txt
