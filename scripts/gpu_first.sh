#!/bin/bash
# First GPU contact: tests, smoke, bench (ours + baseline), launch list, one ncu capture.
set -x
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
python -c "import torch; print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0))" > gpurun_out/device.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -2 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
timeout 900 python bench.py --impl baseline --steps 2 --warmup 3 > gpurun_out/bench_baseline.json 2> gpurun_out/bench_baseline.err; tail -2 gpurun_out/bench_baseline.json; tail -5 gpurun_out/bench_baseline.err
# per-phase timing
timeout 600 python scripts/phase_times.py > gpurun_out/phase_times.txt 2>&1; tail -20 gpurun_out/phase_times.txt
# launch list (cold-cache serialised; shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_bench.log 2>&1
# full capture of the coordinate-select kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:coord_select -c 2 -o gpurun_out/prof_select python scripts/run_select_once.py > gpurun_out/ncu_select.log 2>&1
ls -la gpurun_out
