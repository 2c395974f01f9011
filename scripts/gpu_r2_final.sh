#!/bin/bash
# Final state check of the round on one GPU: full GPU suite, smoke, headline bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_multigpu.py > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_final.log | cut -c1-300
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -1 gpurun_out/bench_final.json | cut -c1-330; tail -1 gpurun_out/bench_final.json | grep -o '"e2e".\{0,200\}'; tail -2 gpurun_out/bench_final.err
