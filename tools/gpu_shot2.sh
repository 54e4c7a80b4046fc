#!/bin/bash
# round-2 GPU shot: tests with the new default (13) and the tile queue (17), variant timing, the bench line, ncu evidence
mkdir -p gpurun_out
timeout -k 5 500 python -m pytest tests -q -m gpu -x > gpurun_out/r02_gpu_tests.log 2>&1
echo "full pytest rc=$?" >> gpurun_out/r02_gpu_tests.log
timeout -k 5 200 python tools/kbench.py 8192 13,17,18,19,20,21,10,13,17,18,19,20,21 > gpurun_out/r02_kbench_17_21.log 2>&1
timeout -k 5 300 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout -k 5 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
bash tools/profile_bench.sh > gpurun_out/r02_profile.log 2>&1
tail -n 4 gpurun_out/r02_gpu_tests.log gpurun_out/r02_kbench_17_21.log gpurun_out/r02_bench_n1.err gpurun_out/r02_profile.log
