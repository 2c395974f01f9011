#!/bin/bash
# Multi-GPU validation (run with gpurun --gpus N): sharded-vs-single tests, then the bench at 1 and N GPUs.
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 1500 python -m pytest tests/test_multigpu.py -m gpu -q -x --timeout 900 > gpurun_out/pytest_multigpu.log 2>&1; echo "exit $?"; tail -25 gpurun_out/pytest_multigpu.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; cat gpurun_out/bench_1gpu.json; tail -3 gpurun_out/bench_1gpu.err
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; cat gpurun_out/bench_${N}gpu.json; tail -15 gpurun_out/bench_${N}gpu.err | cut -c1-300
