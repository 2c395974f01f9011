#!/bin/bash
# Stop experiment sweeps started by scripts/cifar10.sh.  The reference's kill_cifar.sh greps `ps` for "cifar10"
# and kill -9s whatever matches; here the sweep records the PIDs it started in scripts/outputs/pids and only those
# are signalled.
cd "$(dirname "$0")"
[ -f outputs/pids ] || { echo "no recorded runs"; exit 0; }
while read -r pid; do kill "$pid" 2>/dev/null && echo "stopped $pid"; done < outputs/pids
rm -f outputs/pids
