#!/bin/bash
# round-2 GPU shot 5: lanes (frame overlap), conservative beam, full test suite, bench
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -q -m gpu -x > gpurun_out/r02_gpu_tests_b.log 2>&1
echo "full pytest rc=$?" >> gpurun_out/r02_gpu_tests_b.log
timeout -k 5 300 python bench.py --steps 300 > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err
timeout -k 5 300 python bench.py --steps 300 --kernel 17 --no-cpu-baseline > gpurun_out/r02_bench_n1_k17.json 2> gpurun_out/r02_bench_n1_k17.err
timeout -k 5 200 python tools/beam_bench.py 8192 0 > gpurun_out/r02_beam_bench_mode0.log 2>&1
timeout -k 5 200 python tools/beam_bench.py 8192 2 > gpurun_out/r02_beam_bench_mode2.log 2>&1
tail -n 6 gpurun_out/r02_gpu_tests_b.log gpurun_out/r02_bench_n1_c.err gpurun_out/r02_beam_bench_mode0.log
