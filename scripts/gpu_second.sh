#!/bin/bash
# Second GPU contact: validate tcgen05 kernels in isolated processes, then bench + phase profile.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -2 gpurun_out/build.log
for k in gram grouped_wgrad im2col; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$k" --timeout 120 > gpurun_out/pytest_$k.log 2>&1
  echo "pytest $k exit $?"; tail -12 gpurun_out/pytest_$k.log
done
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -2 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
BLADES_WGRAD=cublas timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_ours_cublas_wgrad.json 2> gpurun_out/bench_ours_cublas.err; tail -2 gpurun_out/bench_ours_cublas_wgrad.json; tail -5 gpurun_out/bench_ours_cublas.err
timeout 600 python scripts/phase_times.py > gpurun_out/phase_times.txt 2>&1; head -12 gpurun_out/phase_times.txt; sed -n 12,60p gpurun_out/phase_times.txt | cut -c1-200
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1; cat gpurun_out/kernel_bench.txt
ls -la gpurun_out
