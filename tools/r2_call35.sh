#!/bin/bash
# BASELINE configs C3 / C4 with the shipped library; bandwidth of the pitched head-group copies
mkdir -p gpurun_out
timeout 900 python tools/c3_trajectory.py > gpurun_out/c3_trajectory_r2.json 2> gpurun_out/c35_c3.err; tail -2 gpurun_out/c35_c3.err; python -c "
import json; d=json.loads(open('gpurun_out/c3_trajectory_r2.json').read().strip().splitlines()[-1]); print({k:(v if not isinstance(v,list) else v[:3]+['...']+v[-3:]) for k,v in d.items()})"
timeout 900 python tools/c4_calibrated.py > gpurun_out/c4_calibrated_r2.json 2> gpurun_out/c35_c4.err; tail -2 gpurun_out/c35_c4.err; tail -c 1500 gpurun_out/c4_calibrated_r2.json
timeout 300 python tools/copy2d_rate.py > gpurun_out/copy2d_rate_r2.txt 2>&1; cat gpurun_out/copy2d_rate_r2.txt
