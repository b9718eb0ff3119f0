#!/bin/bash
# final-candidate check: whole GPU suite, update timings, full bench line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/c26_pytest.log 2>&1; tail -4 gpurun_out/c26_pytest.log
timeout 300 python tools/time_update.py > gpurun_out/c26_upd.txt 2>&1; cat gpurun_out/c26_upd.txt
timeout 1500 python bench.py > gpurun_out/c26_bench.json 2> gpurun_out/c26_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c26_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c26_bench.json').read().strip().splitlines()[-1])
print("ms/step",d['ms_per_step'],"value",d['value'],"e2e",d['e2e']['ms_per_step'],d['e2e']['value'])
print(d['e2e']['api']); print(d['roofline']); print(d['gpu_comparators']['summary']); print(d['cpu_baseline']); print(d['clocks'])
PY
