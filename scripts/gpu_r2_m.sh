#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_multigpu.py > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu_m.log | cut -c1-300
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; tail -1 gpurun_out/bench_m.json | cut -c1-330; tail -1 gpurun_out/bench_m.json | grep -o '"e2e".\{0,200\}'; tail -2 gpurun_out/bench_m.err
for g in 1 0; do
BLADES_ROUND_GRAPH=$g timeout 300 python bench.py --config multikrum --steps 12 --warmup 4 --no-port > gpurun_out/bench_multikrum_graph$g.json 2> gpurun_out/bench_multikrum_graph$g.err; echo "multikrum round_graph=$g"; tail -1 gpurun_out/bench_multikrum_graph$g.json | cut -c1-230; tail -1 gpurun_out/bench_multikrum_graph$g.json | grep -o '"e2e".\{0,120\}'; tail -2 gpurun_out/bench_multikrum_graph$g.err
done
timeout 300 python scripts/profile_round.py > gpurun_out/round_kernels_m.txt 2>&1; head -22 gpurun_out/round_kernels_m.txt | cut -c1-120
