#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -2 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
timeout 600 python scripts/phase_times.py > gpurun_out/phase_times.txt 2>&1; head -12 gpurun_out/phase_times.txt; sed -n 12,70p gpurun_out/phase_times.txt | cut -c1-220
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1; cat gpurun_out/kernel_bench.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"coord_select|gram_tcgen05|wgrad_tcgen05|row_combine" -c 6 -o gpurun_out/prof_kernels python scripts/run_kernels_once.py > gpurun_out/ncu_kernels.log 2>&1; tail -3 gpurun_out/ncu_kernels.log
ls -la gpurun_out
