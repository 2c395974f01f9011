#!/bin/bash
# 8 GPUs, short: bench with push-mode aggregation (parallel DMA streams) vs pull, + the aggregation microbenchmark.
set -x
N=${1:-8}
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 30 --warmup 5 --no-port > gpurun_out/bench_${N}gpu_$name.json 2> gpurun_out/bench_${N}gpu_$name.err
  tail -1 gpurun_out/bench_${N}gpu_$name.json | cut -c1-200; tail -1 gpurun_out/bench_${N}gpu_$name.json | grep -o '"e2e".\{0,330\}'; grep -v "OMP\|\*\*\*" gpurun_out/bench_${N}gpu_$name.err | tail -3 | cut -c1-300
}
run push BLADES_X=1
run pull BLADES_AGG_PUSH=0
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 scripts/agg_bench_multigpu.py > gpurun_out/agg_bench_$N.json 2> gpurun_out/agg_bench_$N.err; tail -1 gpurun_out/agg_bench_$N.json | cut -c1-900; grep -v "OMP\|\*\*\*" gpurun_out/agg_bench_$N.err | tail -3 | cut -c1-300
