#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_multigpu.py > gpurun_out/pytest_gpu_p.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_p.log | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; tail -1 gpurun_out/bench_p.json | cut -c1-330; tail -1 gpurun_out/bench_p.json | grep -o '"e2e".\{0,200\}'; tail -2 gpurun_out/bench_p.err
timeout 300 python scripts/profile_round.py > gpurun_out/round_kernels_p.txt 2>&1; head -20 gpurun_out/round_kernels_p.txt | cut -c1-120
