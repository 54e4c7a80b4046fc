#!/bin/bash
# 8 GPUs: strong scaling of one 1080p frame with 4 frames in flight per GPU, lanes ablation, 4K, BASELINE configs[4] (progressive 64 spp at 8192^2)
mkdir -p gpurun_out
run() { # name nproc args...
  name=$1; n=$2; shift 2
  timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
}
run r02_scale_n8 8 bench.py --gpus 8 --steps 400 --warmup 8
run r02_scale_n4 4 bench.py --gpus 4 --steps 400 --warmup 8
run r02_scale_n2 2 bench.py --gpus 2 --steps 400 --warmup 8
timeout -k 5 240 python bench.py --steps 400 --warmup 8 --no-cpu-baseline > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
run r02_scale_n8_lanes2 8 bench.py --gpus 8 --steps 400 --warmup 8 --lanes 2
run r02_scale_n8_lanes3 8 bench.py --gpus 8 --steps 400 --warmup 8 --lanes 3
run r02_scale_n8_4k 8 bench.py --gpus 8 --steps 300 --warmup 8 --width 3840 --height 2160
run r02_scale_n8_frames 8 bench.py --gpus 8 --steps 300 --warmup 8 --partition frames
run r02_config5_n8 8 tools/c5_progressive.py
run r02_config5_n1 1 tools/c5_progressive.py --check-rows 0
tail -n 3 gpurun_out/r02_scale_n8.err gpurun_out/r02_config5_n8.err
