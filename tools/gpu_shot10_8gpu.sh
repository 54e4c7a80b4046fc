#!/bin/bash
# 8 GPUs: frames in flight 4 / 5 / 6 with the fused wait+signal launch, and BASELINE configs[4]
mkdir -p gpurun_out
run() { name=$1; n=$2; shift 2
  timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
}
run r02_scale2_n8_lanes4 8 bench.py --gpus 8 --steps 400 --warmup 8 --lanes 4
run r02_scale2_n8_lanes6 8 bench.py --gpus 8 --steps 400 --warmup 8 --lanes 6
run r02_scale2_n8_lanes5 8 bench.py --gpus 8 --steps 400 --warmup 8 --lanes 5
run r02_scale2_n4_lanes6 4 bench.py --gpus 4 --steps 400 --warmup 8 --lanes 6
run r02_config5_n8 8 tools/c5_progressive.py
run r02_config5_n1 1 tools/c5_progressive.py --check-rows 0
timeout -k 5 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "two_gpu or lanes or interleaved" > gpurun_out/r02_two_gpu_tests_b.log 2>&1
tail -n 3 gpurun_out/r02_two_gpu_tests_b.log gpurun_out/r02_config5_n8.err
