#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_gram_solve_gpu.py -m gpu -q --timeout 300 -k "gram or fltrust or krum or weiszfeld or centered or combine or aggregators" 2>&1 | tail -4
timeout 300 python scripts/kernel_bench.py 2>&1 | grep -E "gram|trimmed|median|row_combine" | tee gpurun_out/kernel_bench_h.txt
timeout 200 python scripts/e2e_timeline.py 2>&1 | tail -8 | cut -c1-200 | tee gpurun_out/e2e_timeline_1gpu.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"coord_select_part|gram_tcgen05" -c 6 -o gpurun_out/prof_select_full python scripts/run_select_full.py --both > gpurun_out/ncu_select_full.log 2>&1; tail -2 gpurun_out/ncu_select_full.log
