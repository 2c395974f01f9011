#!/bin/bash
set -x
N=${1:-4}
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 30 --warmup 5 --no-port > gpurun_out/bench_${N}gpu_final.json 2> gpurun_out/bench_${N}gpu_final.err
tail -1 gpurun_out/bench_${N}gpu_final.json | cut -c1-200; tail -1 gpurun_out/bench_${N}gpu_final.json | grep -o '"e2e".\{0,200\}'; grep -v "OMP\|\*\*\*" gpurun_out/bench_${N}gpu_final.err | tail -3 | cut -c1-300
