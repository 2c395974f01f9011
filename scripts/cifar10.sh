#!/bin/bash
# Attack x aggregator sweep on federated CIFAR-10 (reference scripts/cifar10.sh, made runnable).
# Usage: NGPU=8 DATASET=synthetic-cifar10 bash scripts/cifar10.sh
NGPU=${NGPU:-1}
DATASET=${DATASET:-cifar10}
PIDS="$(dirname "$0")/outputs/pids"       # scripts/stop_runs.sh signals exactly these
mkdir -p "$(dirname "$PIDS")"; echo $$ >> "$PIDS"
for seed in 1; do
  for attack in noise labelflipping signflipping alie ipm; do
    for agg in mean median trimmedmean krum geomed autogm clustering clippedclustering centeredclipping; do
      CMD="scripts/main.py --use-cuda --dataset $DATASET --attack $attack --agg $agg --seed $seed --num_byzantine 8 --global_round 600 --local_round 50"
      if [ "$NGPU" -gt 1 ]; then
        python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NGPU" --master-addr 127.0.0.1 $CMD &
      else
        python $CMD &
      fi
      echo $! >> "$PIDS"; wait $!
    done
  done
done
