#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_multigpu.py > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_o.log | cut -c1-300
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; tail -1 gpurun_out/bench_o.json | cut -c1-330; tail -1 gpurun_out/bench_o.json | grep -o '"e2e".\{0,200\}'; tail -2 gpurun_out/bench_o.err
timeout 300 python scripts/profile_round.py > gpurun_out/round_kernels_o.txt 2>&1; head -22 gpurun_out/round_kernels_o.txt | cut -c1-120
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_smoke_o.csv python __graft_entry__.py smoke > gpurun_out/ncu_smoke_o.log 2>&1; tail -1 gpurun_out/ncu_smoke_o.log
