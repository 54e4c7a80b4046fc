#!/bin/bash
# occupancy variants of the default kernel: 13 (64 regs, 8 CTAs/SM) vs 25 (56 regs, 9) vs 26 (48 regs, 10), same bench command
mkdir -p gpurun_out
for k in 13 25 26 13 25; do
  timeout 200 python bench.py --kernel $k --no-cpu-baseline --steps 300 --warmup 5 > gpurun_out/occ_k$k.json 2> gpurun_out/occ_k$k.err
  echo "kernel $k: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/occ_k$k.json | head -1)"
done
timeout 200 python bench.py --kernel 25 --no-cpu-baseline --beam 0 --lanes 2 --steps 300 --warmup 5 > gpurun_out/occ_k25_nobeam.json 2> /dev/null
echo "kernel 25 nobeam lanes2: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/occ_k25_nobeam.json | head -1)"
timeout 120 python -m pytest tests/test_zz_gpu_variants.py -q -x -k "ctas" 2>&1 | tail -2
