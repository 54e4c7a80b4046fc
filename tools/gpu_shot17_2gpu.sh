#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_gpu_parity.py -q -x -k "two_gpu" > gpurun_out/r02_two_gpu_tests_c.log 2>&1
echo "rc=$?" >> gpurun_out/r02_two_gpu_tests_c.log
tail -n 15 gpurun_out/r02_two_gpu_tests_c.log
