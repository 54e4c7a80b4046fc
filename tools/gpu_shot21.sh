#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 60 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_variants.py -q -x -k "ray_binning or ray_stream or persistent_stream" > gpurun_out/r02_sortkey_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02_sortkey_tests.log
timeout -k 5 60 python tools/sortkey_bench.py 2048 1024 8 > gpurun_out/r02_sortkey_bench.log 2>&1
tail -n 3 gpurun_out/r02_sortkey_tests.log; cat gpurun_out/r02_sortkey_bench.log | tail -6
