#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 300 --warmup 5 > gpurun_out/r02_bench_n2_hostgather.json 2> gpurun_out/r02_bench_n2_hostgather.err
timeout -k 5 300 python bench.py --steps 300 > gpurun_out/r02_bench_n1_e.json 2> gpurun_out/r02_bench_n1_e.err
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "two_gpu or lanes" > gpurun_out/r02_two_gpu_tests_c.log 2>&1
tail -n 5 gpurun_out/r02_bench_n2_hostgather.err gpurun_out/r02_two_gpu_tests_c.log
