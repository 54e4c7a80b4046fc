#!/bin/bash
mkdir -p gpurun_out
for q in "--quick" ""; do
timeout -k 5 300 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python tools/sanitize_gpu.py $q > gpurun_out/r02_sanitizer_racecheck$q.log 2>&1
echo "racecheck $q rc=$?" | tee -a gpurun_out/r02_sanitizer_racecheck$q.log
grep -E "RACECHECK SUMMARY|sanitize sequence ok" gpurun_out/r02_sanitizer_racecheck$q.log | tail -3
done
timeout -k 5 200 compute-sanitizer --tool initcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_gpu.py --quick > gpurun_out/r02_sanitizer_initcheck.log 2>&1
echo "initcheck rc=$?" | tee -a gpurun_out/r02_sanitizer_initcheck.log
grep -E "ERROR SUMMARY|sanitize sequence ok" gpurun_out/r02_sanitizer_initcheck.log | tail -3
