#!/bin/bash
# Multi-GPU validation: sharded-vs-single tests (multimem epilogue, in-switch Gram reduce, pipelined windows), bench A/B.
#   gpurun --gpus N -- 'bash scripts/gpu_r2_mg.sh N'
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 600 ${MGPU_K:+-k "$MGPU_K"} > gpurun_out/pytest_multigpu_$N.log 2>&1; echo "exit $?"; tail -25 gpurun_out/pytest_multigpu_$N.log | cut -c1-300
run() { # name, env...
  name=$1; shift
  env "$@" NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${N}gpu_$name.json 2> gpurun_out/bench_${N}gpu_$name.err
  tail -1 gpurun_out/bench_${N}gpu_$name.json | cut -c1-300; tail -1 gpurun_out/bench_${N}gpu_$name.json | grep -o '"e2e".\{0,260\}'; tail -2 gpurun_out/bench_${N}gpu_$name.err | cut -c1-300
}
run default BLADES_X=1
run nopipe BLADES_AGG_PIPELINE=0
run nomc BLADES_MULTIMEM=0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 scripts/phase_times_multigpu.py > gpurun_out/phase_times_$N.json 2> gpurun_out/phase_times_$N.err; tail -1 gpurun_out/phase_times_$N.json | cut -c1-600; tail -2 gpurun_out/phase_times_$N.err | cut -c1-300
