#!/bin/bash
# 8 GPUs: tile partition with / without the per-rank conservative beam pre-pass; host-gather e2e
mkdir -p gpurun_out
run() { name=$1; n=$2; shift 2
  timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
}
run r02_scale3_n8_beam1 8 bench.py --gpus 8 --steps 400 --warmup 8 --beam 1
run r02_scale3_n8_beam0 8 bench.py --gpus 8 --steps 400 --warmup 8 --beam 0
run r02_scale3_n4_beam1 4 bench.py --gpus 4 --steps 400 --warmup 8 --beam 1
run r02_scale3_n4_beam0 4 bench.py --gpus 4 --steps 400 --warmup 8 --beam 0
run r02_scale3_n2_beam1 2 bench.py --gpus 2 --steps 400 --warmup 8 --beam 1
tail -n 3 gpurun_out/r02_scale3_n8_beam1.err
