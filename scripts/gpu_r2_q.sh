#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_multigpu.py > gpurun_out/pytest_gpu_q.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_q.log | cut -c1-300
for v in 1 0; do
BLADES_BN_REMASK=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-port > gpurun_out/bench_q$v.json 2> gpurun_out/bench_q$v.err; echo "remask=$v"; tail -1 gpurun_out/bench_q$v.json | cut -c1-230; tail -1 gpurun_out/bench_q$v.json | grep -o '"e2e".\{0,120\}'; tail -2 gpurun_out/bench_q$v.err
done
