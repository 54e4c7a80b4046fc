#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -q -m gpu -x -s > gpurun_out/r02_gpu_tests_c.log 2>&1
echo "full pytest rc=$?" >> gpurun_out/r02_gpu_tests_c.log
for l in 1 2 3 4; do
  timeout -k 5 200 python bench.py --steps 300 --lanes $l --no-cpu-baseline > gpurun_out/r02_bench_n1_lanes$l.json 2> gpurun_out/r02_bench_n1_lanes$l.err
done
grep -h "8192^3" gpurun_out/r02_gpu_tests_c.log | head; tail -n 5 gpurun_out/r02_gpu_tests_c.log
