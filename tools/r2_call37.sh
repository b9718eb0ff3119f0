#!/bin/bash
# final evidence set of the shipped library: whole GPU suite, default bench line, ncu launch list + captures, update timings
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/c37_pytest.log 2>&1; tail -3 gpurun_out/c37_pytest.log
timeout 1500 python bench.py > gpurun_out/c37_bench.json 2> gpurun_out/c37_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c37_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c37_bench_ref.json 2>> gpurun_out/c37_bench.err; echo "ref rc=$?"; cat gpurun_out/c37_bench_ref.json | cut -c1-400
timeout 300 python tools/time_update.py > gpurun_out/c37_upd.txt 2>&1
timeout 1500 bash tools/capture_profiles.sh r2 > gpurun_out/c37_capture.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:la_fwd_kernel -s 1 -c 1 -o gpurun_out/prof_fwd_wan00_r2 -f python tools/one_launch.py --sparsity 0 > gpurun_out/prof_fwd_wan00_r2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c37_bench.json').read().strip().splitlines()[-1])
print("ms/step",d['ms_per_step'],"value",d['value'],"e2e",d['e2e']['ms_per_step'],d['e2e']['value'])
print(d['roofline']); print(d['gpu_comparators']['summary']); print(d['clocks'])
PY
