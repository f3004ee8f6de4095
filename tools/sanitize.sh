#!/bin/bash
# compute-sanitizer over every kernel family (run under gpurun). Writes gpurun_out/<tag>_sanitize_<tool>.txt
R=${1:-r01}
O=gpurun_out
mkdir -p $O
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --launch-timeout 300 --print-limit 20 python tools/sanitize_driver.py > $O/${R}_sanitize_${tool}.txt 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize driver ok' $O/${R}_sanitize_${tool}.txt | tr '\n' ' ')"
done
