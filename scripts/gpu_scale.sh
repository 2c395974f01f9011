#!/bin/bash
# Scaling run on N GPUs of one box: bench at each listed size + a quick sharded-vs-single check.
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for n in "$@"; do
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
  else
    NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  tail -1 gpurun_out/scale_$n.json | cut -c1-220; tail -1 gpurun_out/scale_$n.json | grep -o '"e2e".*'; grep -E "Error|error" gpurun_out/scale_$n.err | head -5
done
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x --timeout 600 -k "trimmedmean-alie-mlp or geomed" > gpurun_out/pytest_multigpu.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_multigpu.log | cut -c1-300
