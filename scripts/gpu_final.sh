#!/bin/bash
# Validation + evidence run on one GPU: all GPU tests, smoke, bench, kernel bench, sanitizer, ncu captures.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_multigpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -1 gpurun_out/bench_ours.json | cut -c1-1500
timeout 400 python scripts/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1; cat gpurun_out/kernel_bench.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 500 -k "trimmed_mean_and_median or row_combine or fill_normal or attack_rows or im2col_nhwc or client_bn_nhwc or wgrad_padded or gram or gather_samples or conv_wgrad_implicit" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/sanitizer_memcheck.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"coord_select|gram_tcgen05|wgrad_tcgen05|row_combine" -c 6 -o gpurun_out/prof_kernels python scripts/run_kernels_once.py > gpurun_out/ncu_kernels.log 2>&1; tail -2 gpurun_out/ncu_kernels.log
# reference arm (unmodified reference through its own API; one round is minutes of CPU aggregation)
BLADES_REF_BUDGET_S=${REF_BUDGET:-60} BLADES_REF_DEADLINE_S=${REF_DEADLINE:-420} timeout 460 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -1 gpurun_out/bench_reference.json | cut -c1-600
ls -la gpurun_out | head -30
