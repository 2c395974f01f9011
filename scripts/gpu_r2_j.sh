#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "select or trimmed or median or partition" 2>&1 | tail -3
timeout 200 python scripts/kernel_bench.py 2>&1 | grep -E "trimmed|median" | tee gpurun_out/kernel_bench_j.txt
BLADES_SELECT_BLOCK=256 timeout 200 python scripts/kernel_bench.py 2>&1 | grep -E "ALIE" | sed "s/^/block=256 /" | tee -a gpurun_out/kernel_bench_j.txt
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; tail -1 gpurun_out/bench_j.json | cut -c1-250;  tail -1 gpurun_out/bench_j.json | grep -o '"e2e".\{0,330\}'; tail -2 gpurun_out/bench_j.err
