#!/bin/bash
# 8 GPUs: strong scaling of ONE 1080p frame over the tile partition (N = 2, 4, 8), and the 4K frame at N = 8
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_8gpu.log 2>&1
for n in 8 4 2; do
  timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 300 --warmup 5 > gpurun_out/r02_bench_n${n}_tiles.json 2> gpurun_out/r02_bench_n${n}_tiles.err
done
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 300 --warmup 5 --width 3840 --height 2160 > gpurun_out/r02_bench_n8_tiles_4k.json 2> gpurun_out/r02_bench_n8_tiles_4k.err
timeout -k 5 240 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n1_same_box.json 2> gpurun_out/r02_bench_n1_same_box.err
timeout -k 5 240 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --width 3840 --height 2160 > gpurun_out/r02_bench_n1_4k.json 2> gpurun_out/r02_bench_n1_4k.err
tail -n 3 gpurun_out/r02_bench_n8_tiles.err
