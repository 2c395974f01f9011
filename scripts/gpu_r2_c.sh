#!/bin/bash
# Select kernels without the per-value sanitise (slow path on non-finite totals), CE mixes 2/3/4, device Gram solvers.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_multigpu.py > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu_c.log | cut -c1-300
for v in 2 3 4 0; do
  BLADES_SELECT_CE=$v timeout 200 python scripts/kernel_bench.py 2>&1 | grep -E "trimmed_mean|median" | sed "s/^/CE=$v  /" | tee -a gpurun_out/kernel_bench_ce2.txt
done
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours_c.json 2> gpurun_out/bench_ours_c.err; tail -1 gpurun_out/bench_ours_c.json | cut -c1-400
