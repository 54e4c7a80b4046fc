#!/bin/bash
# 1 GPU: incremental upload timing, BASELINE configs 2-4 at full size (kept JSON / logs), ray-stream launch list (sort share)
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "incremental" > gpurun_out/r02_incremental_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_incremental_tests.log
timeout -k 5 300 python bench.py --size 2048 --steps 300 > gpurun_out/r02_config2_2048.json 2> gpurun_out/r02_config2_2048.err
timeout -k 5 400 python bench.py --width 3840 --height 2160 --casts 5 --mirror 3 --steps 60 --cpu-seconds 20 > gpurun_out/r02_config3_4k_5casts_mirror.json 2> gpurun_out/r02_config3_4k_5casts_mirror.err
timeout -k 5 300 python tools/stream_bench.py 8192 67108864 > gpurun_out/r02_config4_64Mi_rays.log 2>&1
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_stream_launches.csv python tools/stream_bench.py 8192 16777216 > gpurun_out/r02_stream_under_ncu.log 2>&1
timeout -k 5 300 python bench.py --steps 300 > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err
grep -h "8192^3" gpurun_out/r02_incremental_tests.log; tail -n 3 gpurun_out/r02_incremental_tests.log gpurun_out/r02_config4_64Mi_rays.log
