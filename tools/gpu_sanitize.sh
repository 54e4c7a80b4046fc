#!/bin/bash
# compute-sanitizer over one pass of every kernel family (tools/sanitize_gpu.py); summaries -> gpurun_out/r02_sanitizer_*.log
mkdir -p gpurun_out
timeout 120 python tools/sanitize_gpu.py > gpurun_out/r02_sanitizer_plain.log 2>&1
echo "plain rc=$?" | tee -a gpurun_out/r02_sanitizer_plain.log
for tool in memcheck racecheck synccheck; do
  extra=""; [ $tool = racecheck ] && extra="--quick"
  timeout -k 5 330 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_gpu.py $extra > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r02_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize sequence ok" gpurun_out/r02_sanitizer_$tool.log | tail -3
done
