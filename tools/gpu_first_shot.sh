#!/bin/bash
# One gpurun call that answers the open questions of DESIGN.md section 7 (about 5 minutes on one B200):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/gpu_first_shot.sh'
# 1. device parity of the kernel variants, including the not-yet-measured variant 13 (inline PTX)
# 2. per-camera kernel times: default (10) against 13, 0 and the FFMA build
# 3. if 13 is bit-exact: ncu --set full of the three bench frames under 10 and 13 (pipe utilisation, stall reasons)
# 4. the ray-stream kernels (grid-stride against persistent threads with warp-level ray fetch) on 16 Mi incoherent rays
# 5. the bench line and its ncu launch list with the library default
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests -q -m gpu -x > gpurun_out/r02_gpu_tests_first.log 2>&1
echo "full pytest rc=$?" >> gpurun_out/r02_gpu_tests_first.log
timeout -k 5 200 python -m pytest tests/test_zz_gpu_variants.py -q -m gpu > gpurun_out/variants_test.log 2>&1
echo "pytest rc=$?" >> gpurun_out/variants_test.log
timeout -k 5 120 python tools/kbench.py 8192 10,13,15,16,14,0,10,13,15,16,0f > gpurun_out/kbench_13.log 2>&1
if grep -q "rc=0" gpurun_out/variants_test.log; then
  timeout -k 5 150 ncu --set full --clock-control none -k regex:k_render_tile -c 6 -f -o gpurun_out/tile_10_13 python tools/ncu_ab.py 8192 ABC 10,13 > gpurun_out/ncu_10_13.log 2>&1
fi
timeout -k 5 120 python tools/stream_bench.py 8192 16777216 > gpurun_out/stream_bench.log 2>&1  # grid-stride vs persistent ray-stream kernel
timeout -k 5 180 python bench.py --steps 300 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout -k 5 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -n 6 gpurun_out/r02_gpu_tests_first.log gpurun_out/variants_test.log gpurun_out/kbench_13.log gpurun_out/stream_bench.log
