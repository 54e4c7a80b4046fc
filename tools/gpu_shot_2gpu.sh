#!/bin/bash
# 2 GPUs: the tile partition over NVLink (test + bench), tiles vs frames
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_2gpu.log 2>&1
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "two_gpu or interleaved or fence" > gpurun_out/r02_two_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_two_gpu_tests.log
for part in tiles frames; do
  timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 5 --partition $part > gpurun_out/r02_bench_n2_$part.json 2> gpurun_out/r02_bench_n2_$part.err
done
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 300 --warmup 5 --kernel 0 > gpurun_out/r02_bench_n2_tiles_k0.json 2> gpurun_out/r02_bench_n2_tiles_k0.err
tail -n 5 gpurun_out/r02_two_gpu_tests.log gpurun_out/r02_bench_n2_tiles.err
