#!/bin/bash
set -x
mkdir -p gpurun_out
SAN_TIMEOUT=400 bash scripts/sanitize.sh memcheck
for t in 1024 4096; do
BLADES_IMPLICIT_MAX_T=$t timeout 300 python bench.py --steps 30 --warmup 5 --no-port --no-e2e > gpurun_out/bench_l_$t.json 2> gpurun_out/bench_l_$t.err; echo "implicit_max_t=$t"; tail -1 gpurun_out/bench_l_$t.json | cut -c1-230; tail -2 gpurun_out/bench_l_$t.err
done
