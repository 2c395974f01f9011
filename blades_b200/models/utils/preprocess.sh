#!/usr/bin/env bash
# LEAF-style preprocessing pipeline (reference models/utils/preprocess.sh):
#   preprocess.sh --name DIR [-s iid|niid] [--sf FRAC] [-k MIN_SAMPLES] [-t user|sample] [--tf TRAIN_FRAC] [--seed S]
NAME=""; SAMPLE="niid"; SFRAC="0.1"; MINS="10"; SPLIT="sample"; TFRAC="0.9"; SEED=""
while [[ $# -gt 0 ]]; do case $1 in
  --name) NAME="$2"; shift 2;; -s) SAMPLE="$2"; shift 2;; --sf) SFRAC="$2"; shift 2;; -k) MINS="$2"; shift 2;;
  -t) SPLIT="$2"; shift 2;; --tf) TFRAC="$2"; shift 2;; --seed) SEED="--seed $2"; shift 2;; *) shift;; esac; done
[ -z "$NAME" ] && { echo "--name required"; exit 1; }
M=blades_b200.models.utils
python -m $M.sample --name "$NAME" --$SAMPLE --fraction "$SFRAC" $SEED
python -m $M.remove_users --name "$NAME" --min_samples "$MINS"
if [ "$SPLIT" = "user" ]; then BY="--by_user"; else BY="--by_sample"; fi
python -m $M.split_data --name "$NAME" $BY --frac "$TFRAC" $SEED
python -m $M.stats --name "$NAME"
