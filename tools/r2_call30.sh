#!/bin/bash
# tagged-slot exchange as the default: whole GPU suite, hang hunt (fresh processes, back-to-back launches), bench, cycle accounting
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/c30_pytest.log 2>&1; tail -3 gpurun_out/c30_pytest.log
ok=0; bad=0
for i in $(seq 1 40); do
  if timeout 120 python tools/flaky2.py > gpurun_out/c30_flaky_last.txt 2>&1; then ok=$((ok+1)); else bad=$((bad+1)); cp gpurun_out/c30_flaky_last.txt gpurun_out/c30_flaky_bad_$i.txt; fi
done
echo "flaky2: ok=$ok bad=$bad" | tee gpurun_out/c30_flaky.txt; tail -2 gpurun_out/c30_flaky_last.txt
timeout 900 python tools/prof_clocks.py > gpurun_out/c30_prof_clocks.txt 2>&1; tail -32 gpurun_out/c30_prof_clocks.txt
timeout 1500 python bench.py > gpurun_out/c30_bench.json 2> gpurun_out/c30_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c30_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c30_bench.json').read().strip().splitlines()[-1])
print("ms/step",d['ms_per_step'],"value",d['value'],"e2e",d['e2e']['ms_per_step'],d['e2e']['value'])
print(d['roofline']); print(d['gpu_comparators']['summary']); print(d['sweep'])
PY
