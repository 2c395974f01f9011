#!/bin/bash
# Round-2 state check on one GPU: full GPU suite, smoke, bench, per-kernel round profile, kernel bench,
# launch list under ncu, one full ncu capture of the top kernels.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_r2_a.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv | head -3
timeout 700 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_multigpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -1 gpurun_out/bench_ours.json | cut -c1-1800
timeout 200 python scripts/profile_round.py > gpurun_out/round_kernels.txt 2>&1; head -45 gpurun_out/round_kernels.txt | cut -c1-150
timeout 400 python scripts/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1; cat gpurun_out/kernel_bench.txt | cut -c1-200
# launch list of the smoke run (what the driver records)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_smoke.csv python __graft_entry__.py smoke > gpurun_out/ncu_smoke.log 2>&1; tail -1 gpurun_out/ncu_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"coord_select|gram_tcgen05|wgrad_tcgen05|row_combine|conv_tcgen05" -c 8 -o gpurun_out/prof_kernels python scripts/run_kernels_once.py > gpurun_out/ncu_kernels.log 2>&1; tail -2 gpurun_out/ncu_kernels.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tcgen05" -s 2 -c 2 -o gpurun_out/prof_conv python scripts/run_conv_once.py > gpurun_out/ncu_conv.log 2>&1; tail -2 gpurun_out/ncu_conv.log
ls -la gpurun_out | head -30
