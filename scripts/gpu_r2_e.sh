#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_multigpu.py > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu_e.log | cut -c1-300
for v in 1 0; do
BLADES_AGG_PIPELINE=$v timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_pipe$v.json 2> gpurun_out/bench_pipe$v.err; tail -1 gpurun_out/bench_pipe$v.json | cut -c1-330; tail -1 gpurun_out/bench_pipe$v.json | grep -o '"e2e".\{0,200\}'; tail -3 gpurun_out/bench_pipe$v.err
done
timeout 200 python scripts/kernel_bench.py 2>&1 | grep -E "trimmed|median" | tee gpurun_out/kernel_bench_e.txt
