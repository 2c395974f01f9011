#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "gather or prefetch or select or partition" 2>&1 | tail -3
for v in 148 64 296; do
BLADES_GATHER_CTAS=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-port > gpurun_out/bench_k_$v.json 2> gpurun_out/bench_k_$v.err; echo "ctas=$v"; tail -1 gpurun_out/bench_k_$v.json | cut -c1-230;  tail -1 gpurun_out/bench_k_$v.json | grep -o '"e2e".\{0,200\}'; tail -2 gpurun_out/bench_k_$v.err
done
timeout 200 python scripts/kernel_bench.py 2>&1 | grep -E "trimmed|median" | tee gpurun_out/kernel_bench_k.txt
