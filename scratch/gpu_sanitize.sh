#!/bin/bash
mkdir -p gpurun_out
for tool in ${1:-memcheck racecheck}; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scratch/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported|Invalid|sanitize pass done" gpurun_out/sanitize_$tool.log | awk '!seen[$0]++' | head -12
done
