#!/bin/bash
# Compare-exchange formulation experiment (IMAD-form hi on the FMA pipe) + the select kernels built with it.
set -x
mkdir -p gpurun_out
timeout 120 ./build/ce_bench 2>&1 | tee gpurun_out/ce_bench.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "select or trimmed or median or partition or row_combine" 2>&1 | tail -5
for v in 1 0; do
  BLADES_SELECT_CE=$v timeout 200 python scripts/kernel_bench.py 2>&1 | grep -E "trimmed_mean|median" | sed "s/^/CE=$v  /" | tee -a gpurun_out/kernel_bench_ce.txt
done
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours_b.json 2> gpurun_out/bench_ours_b.err; tail -1 gpurun_out/bench_ours_b.json | cut -c1-400
