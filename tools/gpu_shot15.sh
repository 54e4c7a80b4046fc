#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_gpu_tests_e.log 2>&1
echo "full pytest rc=$?" >> gpurun_out/r02_gpu_tests_e.log
timeout -k 5 300 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
timeout -k 5 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_final_k20.json 2> gpurun_out/r02_bench_final_k20.err
timeout -k 5 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_final_ref.json 2> gpurun_out/r02_bench_final_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
tail -n 4 gpurun_out/r02_gpu_tests_e.log gpurun_out/r02_smoke.log gpurun_out/r02_bench_final.err
