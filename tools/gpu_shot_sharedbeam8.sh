#!/bin/bash
# shared-lattice beam in the tile partition on one 8-GPU box: N = 8 and 4 at 1080p, N = 8 at 4K
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -k "beam_lattice_in_row_parts" > gpurun_out/sb_test.log 2>&1
run() { # N beam extra-tag extra-args
  N=$1; b=$2; tag=$3; shift 3
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$b bench.py --gpus $N --steps 300 --warmup 5 --beam $b "$@" > gpurun_out/sb_n${N}_beam$b$tag.log 2> gpurun_out/sb_n${N}_beam$b$tag.err
  echo "N=$N beam=$b $tag"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/sb_n${N}_beam$b$tag.log; grep -o '"rgba8_mismatch": [0-9]*, "depth_mismatch": [0-9]*' gpurun_out/sb_n${N}_beam$b$tag.log | head -1
}
run 8 2 ""
run 8 0 ""
run 4 2 ""
run 4 0 ""
run 8 2 _4k --width 3840 --height 2160
tail -2 gpurun_out/sb_test.log
