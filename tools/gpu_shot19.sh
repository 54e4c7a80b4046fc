#!/bin/bash
# round-2 closing check after the builder/brush pin: every GPU test, the default bench line, smoke()
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -q -m gpu -x > gpurun_out/r02_gpu_tests_g.log 2>&1
echo "full pytest rc=$?" >> gpurun_out/r02_gpu_tests_g.log
timeout -k 5 200 python bench.py > gpurun_out/r02_bench_final3.json 2> gpurun_out/r02_bench_final3.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke3.log 2>&1
tail -n 4 gpurun_out/r02_gpu_tests_g.log gpurun_out/r02_smoke3.log gpurun_out/r02_bench_final3.err
