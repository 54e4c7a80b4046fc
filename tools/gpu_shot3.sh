#!/bin/bash
# round-2 GPU shot 3: device world builder, register-limit ablations, bench with the device-built world
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_gpu_build.py -q -m gpu -x -s > gpurun_out/r02_gpu_build_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gpu_build_tests.log
timeout -k 5 200 python tools/kbench.py 8192 13,22,23,24,20,13,22,23,24,20 > gpurun_out/r02_kbench_22_24.log 2>&1
timeout -k 5 300 python bench.py --steps 150 > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err
tail -n 6 gpurun_out/r02_gpu_build_tests.log gpurun_out/r02_kbench_22_24.log gpurun_out/r02_bench_n1_b.err
