#!/bin/bash
# tile partition at N = 8, 1080p: band rotation on / off, back to back
mkdir -p gpurun_out
run() { # rotate
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$1 bench.py --gpus 8 --steps 400 --warmup 5 --rotate $1 > gpurun_out/rot_n8_r$1.log 2> gpurun_out/rot_n8_r$1.err
  echo "rotate=$1"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/rot_n8_r$1.log; grep -o '"rgba8_mismatch": [0-9]*, "depth_mismatch": [0-9]*' gpurun_out/rot_n8_r$1.log | head -1
}
run 1
run 0
