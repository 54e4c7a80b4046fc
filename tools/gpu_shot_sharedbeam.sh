#!/bin/bash
# shared-lattice beam in the tile partition: $1 = ranks
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "beam_lattice_in_row_parts or conservative_beam" > gpurun_out/sb_test.log 2>&1
for b in 0 2 1; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$b bench.py --gpus $N --steps 300 --warmup 5 --beam $b > gpurun_out/sb_n${N}_beam$b.log 2> gpurun_out/sb_n${N}_beam$b.err
done
tail -3 gpurun_out/sb_test.log
for b in 0 2 1; do grep -o '"ms_per_step": [0-9.]*' gpurun_out/sb_n${N}_beam$b.log; grep -o '"parity": {[^}]*}' gpurun_out/sb_n${N}_beam$b.log | head -2; done
