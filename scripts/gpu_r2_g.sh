#!/bin/bash
# 2 GPUs: push-mode aggregation (copy-engine DMA of finished windows + local select) vs pull mode.
set -x
N=${1:-2}
mkdir -p gpurun_out
BLADES_MGPU_SIZES=1,$N timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_multigpu_push_$N.log 2>&1; echo "exit $?"; tail -25 gpurun_out/pytest_multigpu_push_$N.log | cut -c1-300
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 30 --warmup 5 --no-port > gpurun_out/bench_${N}gpu_$name.json 2> gpurun_out/bench_${N}gpu_$name.err
  tail -1 gpurun_out/bench_${N}gpu_$name.json | cut -c1-200; tail -1 gpurun_out/bench_${N}gpu_$name.json | grep -o '"e2e".\{0,420\}'; grep -v "OMP\|\*\*\*" gpurun_out/bench_${N}gpu_$name.err | tail -3 | cut -c1-300
}
run push BLADES_X=1
run pull BLADES_AGG_PUSH=0
run push_nopipe BLADES_AGG_PIPELINE=0
