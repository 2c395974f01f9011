#!/bin/bash
set -x
mkdir -p gpurun_out
for v in 2 5 6 1; do
  BLADES_SELECT_CE=$v timeout 200 python scripts/kernel_bench.py 2>&1 | grep -E "ALIE" | sed "s/^/CE=$v  /" | tee -a gpurun_out/kernel_bench_ce3.txt
done
for b in 128 192 320 384 512; do
  BLADES_SELECT_BLOCK=$b timeout 200 python scripts/kernel_bench.py 2>&1 | grep -E "ALIE" | sed "s/^/block=$b  /" | tee -a gpurun_out/kernel_bench_ce3.txt
done
timeout 300 python -m pytest tests/test_gram_solve_gpu.py -m gpu -q 2>&1 | tail -3
