#!/bin/bash
# compute-sanitizer over every kernel of the shipped library (tools/sanitize_small.py)
mkdir -p gpurun_out; : > gpurun_out/sanitizer_r2.txt
timeout 300 python tools/sanitize_small.py >> gpurun_out/sanitizer_r2.txt 2>&1
for tool in memcheck initcheck synccheck racecheck; do
  echo "== compute-sanitizer --tool $tool" >> gpurun_out/sanitizer_r2.txt
  S=900 H=2 timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_small.py 2>&1 | grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|checksum|Error|hazard|Hazard|=========     at' | head -30 >> gpurun_out/sanitizer_r2.txt
done
cat gpurun_out/sanitizer_r2.txt
