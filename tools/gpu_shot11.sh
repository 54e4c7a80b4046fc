#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_gpu_tests_d.log 2>&1
echo "full pytest rc=$?" >> gpurun_out/r02_gpu_tests_d.log
timeout -k 5 300 python bench.py --steps 300 > gpurun_out/r02_bench_n1_beam.json 2> gpurun_out/r02_bench_n1_beam.err
timeout -k 5 300 python bench.py --steps 300 --beam 0 --no-cpu-baseline > gpurun_out/r02_bench_n1_nobeam.json 2> gpurun_out/r02_bench_n1_nobeam.err
timeout -k 5 300 python bench.py --steps 300 --lanes 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_beam_lanes3.json 2> gpurun_out/r02_bench_n1_beam_lanes3.err
timeout -k 5 300 python bench.py --steps 300 --lanes 4 --no-cpu-baseline > gpurun_out/r02_bench_n1_beam_lanes4.json 2> gpurun_out/r02_bench_n1_beam_lanes4.err
bash tools/profile_bench.sh > gpurun_out/r02_profile_b.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
tail -n 4 gpurun_out/r02_gpu_tests_d.log gpurun_out/r02_smoke.log
