#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_multigpu.py -x > gpurun_out/pytest_gpu_f.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu_f.log | cut -c1-300
for v in 1 0; do
BLADES_SELECT_STAGED=$v timeout 200 python scripts/kernel_bench.py 2>&1 | grep -E "trimmed|median" | sed "s/^/staged=$v  /" | tee -a gpurun_out/kernel_bench_f.txt
done
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; tail -1 gpurun_out/bench_f.json | cut -c1-330; tail -2 gpurun_out/bench_f.err
