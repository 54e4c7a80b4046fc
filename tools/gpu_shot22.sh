#!/bin/bash
mkdir -p gpurun_out
timeout -k 3 30 python tools/sortkey_bench.py 2048 1024 4 > gpurun_out/r02_sortkey_bench.log 2>&1
tail -n 6 gpurun_out/r02_sortkey_bench.log
