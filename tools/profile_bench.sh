#!/bin/bash
# The ncu evidence bench.py's roofline.issue / lanes / traffic / gather come from, made from the SAME command the bench
# runs (one GPU; never a multi-rank command):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/profile_bench.sh'
# 1. launch list of a short bench run (per-launch durations; cold-cache, serialised: the kernel's SHARE of the step is
#    what must agree with the bench, not the absolute) -> profiles/r02_launches.csv
# 2. ncu --set full of the render kernel's three camera frames (A, B, C) inside that command -> gpurun_out/r02_bench_kernel.ncu-rep,
#    summarised into profiles/r02_bench_kernel.{json,md} (tools/ncu_summary.py).  bench.py reads the JSON.
# 3. the loop's SASS (cuobjdump) -> profiles/r02_loop_sass.txt
set -u
mkdir -p gpurun_out profiles
K=${SVO_PROFILE_KERNEL:-k_render_tile}
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
# skip the instrumented launches (k_render_stats) and the warm-up: profile three consecutive timed frames
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 3 -f -o gpurun_out/r02_bench_kernel \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_bench_kernel.ncu-rep gpurun_out/r02_bench_kernel > /dev/null 2>&1
cp gpurun_out/r02_bench_kernel.json gpurun_out/r02_bench_kernel.md profiles/ 2>/dev/null
cp gpurun_out/r02_launches.csv profiles/r02_launches.csv 2>/dev/null
tail -n 5 gpurun_out/r02_bench_kernel.md
