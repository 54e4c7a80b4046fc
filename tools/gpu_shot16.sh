#!/bin/bash
# round-2 closing check on one B200: every GPU test, the default bench line, the reference arm, smoke()
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_gpu_tests_f.log 2>&1
echo "full pytest rc=$?" >> gpurun_out/r02_gpu_tests_f.log
timeout -k 5 300 python bench.py > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err
timeout -k 5 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_final2_ref.json 2> gpurun_out/r02_bench_final2_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke2.log 2>&1
tail -n 4 gpurun_out/r02_gpu_tests_f.log gpurun_out/r02_smoke2.log gpurun_out/r02_bench_final2.err
