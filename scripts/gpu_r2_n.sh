#!/bin/bash
# 4 GPUs: headline push vs pull, Multi-Krum (config #4) and GeoMed / ResNet-50 / 512 clients (config #5).
set -x
N=${1:-4}
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --no-port $EXTRA > gpurun_out/bench_${N}gpu_$name.json 2> gpurun_out/bench_${N}gpu_$name.err
  tail -1 gpurun_out/bench_${N}gpu_$name.json | cut -c1-200; tail -1 gpurun_out/bench_${N}gpu_$name.json | grep -o '"e2e".\{0,200\}'; grep -v "OMP\|\*\*\*" gpurun_out/bench_${N}gpu_$name.err | tail -3 | cut -c1-300
}
EXTRA="--steps 30 --warmup 5" run push BLADES_AGG_PUSH=1
EXTRA="--steps 30 --warmup 5" run pull BLADES_AGG_PUSH=0
EXTRA="--config multikrum --steps 12 --warmup 4 --no-e2e" run multikrum BLADES_X=1
EXTRA="--config geomed_r50 --steps 5 --warmup 3 --no-e2e" run geomed_r50 BLADES_X=1
