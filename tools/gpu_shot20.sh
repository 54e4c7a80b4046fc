#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 100 python -m pytest tests/test_gpu_build.py -q -x -k "reference_builders" > gpurun_out/r02_gpu_build_golden.log 2>&1
echo "rc=$?" >> gpurun_out/r02_gpu_build_golden.log
tail -n 5 gpurun_out/r02_gpu_build_golden.log
