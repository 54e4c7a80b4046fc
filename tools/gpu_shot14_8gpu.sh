#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; n=$2; shift 2
  timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
}
run r02_scale4_n8_lanes7 8 bench.py --gpus 8 --steps 400 --warmup 8 --lanes 7
run r02_scale4_n8_lanes6 8 bench.py --gpus 8 --steps 400 --warmup 8 --lanes 6
run r02_scale4_n8_4k 8 bench.py --gpus 8 --steps 300 --warmup 8 --width 3840 --height 2160
tail -n 3 gpurun_out/r02_scale4_n8_lanes7.err
