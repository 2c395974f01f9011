#!/bin/bash
# compute-sanitizer over the kernel numerics tests (needs a GPU; slow -- run under gpurun):
#   memcheck  : every kernel test (select, combine, attack, im2col, BatchNorm incl. cluster/DSMEM, gram, wgrad explicit +
#               implicit, conv fprop/dgrad, pooling / CE / colsum, device Gram solvers)
#   racecheck : shared-memory hazards of the tcgen05 / TMA / mbarrier kernels and the cluster BatchNorm
#   synccheck : barrier usage (bar.sync / mbarrier / cluster barriers) of the same kernels
#   bash scripts/sanitize.sh [memcheck|racecheck|synccheck|all]
TOOL=${1:-all}
mkdir -p gpurun_out
KERNEL_TESTS="tests/test_kernels_gpu.py tests/test_fused_gpu.py tests/test_gram_solve_gpu.py"
# whole simulations replay CUDA graphs and take minutes under the sanitizer: the kernels they launch are covered above
SKIP="not zero_copy and not graph and not simulator and not whole_round and not fedavg and not checkpoint and not prefetch and not short_tail and not evaluation and not pipelined and not schedule and not launches_only and not batched_engine"
TC=${SAN_TC:-"gram or wgrad or conv_tcgen05 or linear_tcgen05 or client_bn or trimmed or median or partition or device_krum or device_weiszfeld or device_centered"}
rc=0
run() { # tool, -k expression
  echo "=== compute-sanitizer --tool $1 -k '$2'"
  timeout ${SAN_TIMEOUT:-1500} compute-sanitizer --tool "$1" --error-exitcode 1 python -m pytest $KERNEL_TESTS -m gpu -q -x --timeout ${SAN_TIMEOUT:-1400} \
      -p no:cacheprovider -k "$2" > gpurun_out/sanitizer_$1.log 2>&1
  r=$?; echo "$1 exit $r"; tail -4 gpurun_out/sanitizer_$1.log | cut -c1-200; grep -c "ERROR SUMMARY" gpurun_out/sanitizer_$1.log
  grep "ERROR SUMMARY" gpurun_out/sanitizer_$1.log | sort | uniq -c | head -5
  [ $r -ne 0 ] && rc=$r
}
case "$TOOL" in
  memcheck) run memcheck "$SKIP" ;;
  racecheck) run racecheck "($TC) and $SKIP" ;;
  synccheck) run synccheck "($TC) and $SKIP" ;;
  all) run memcheck "$SKIP"; run racecheck "($TC) and $SKIP"; run synccheck "($TC) and $SKIP" ;;
esac
exit $rc
