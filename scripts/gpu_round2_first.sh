#!/bin/bash
# First GPU call of the next round (1 GPU, ~8 min): everything that was changed after the last GPU session of round 1.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh'
set -x
mkdir -p gpurun_out
# 1. the whole GPU suite (new: partition select kernel, chunked prefetch, tail batches, capture policy)
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_multigpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
# 2. partition-only trimmed mean vs the full network (separate processes: the switch is read once)
for v in 1 0; do
  BLADES_SELECT_PARTITION=$v timeout 200 python scripts/kernel_bench.py 2>&1 | grep -i "trimmed_mean" | sed "s/^/partition=$v  /" | tee -a gpurun_out/kernel_bench_partition.txt
  BLADES_SELECT_PARTITION=$v timeout 200 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_partition_$v.json 2> gpurun_out/bench_partition_$v.err
  tail -1 gpurun_out/bench_partition_$v.json | cut -c1-130; tail -1 gpurun_out/bench_partition_$v.json | grep -o '"e2e".*' | cut -c1-120
done
# 3. where the round goes now
timeout 200 python scripts/profile_round.py > gpurun_out/round_kernels.txt 2>&1; head -24 gpurun_out/round_kernels.txt | cut -c1-150
# 4. one ncu capture of the new select kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"coord_select_part" -c 1 -o gpurun_out/prof_select_part python scripts/run_kernels_once.py > gpurun_out/ncu_select_part.log 2>&1; tail -2 gpurun_out/ncu_select_part.log
# (multi-GPU follow-up, separate call:  gpurun --gpus 8 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8
#   --master-addr 127.0.0.1 scripts/phase_times_multigpu.py; python -m torch.distributed.run ... bench.py --gpus 8')
