#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c19_pytest.log
tail -6 gpurun_out/c19_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/c19_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c19_bench.json'))
print("ms/step",d['ms_per_step'],"e2e",d['e2e']['ms_per_step'] if d.get('e2e') else None)
r=d['roofline']; print("roofline",r['frac'],r['kernel_ms'],"traffic",r['traffic'],r.get('traffic_source'))
print("update",r['update_kernel'])
for s in d['sweep']: print(s)
print(json.dumps(d.get('gpu_comparators'),indent=1)[:1500])
print(d['e2e'].get('host_link') if d.get('e2e') else None)
PY
