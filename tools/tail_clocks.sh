#!/bin/bash
# Development: clock64() timelines of the tail launches of a few configs (NVPYR_TAIL_DEBUG_CLOCKS, see nvpyr_api.cu).
for c in ${CONFIGS:-"16384" "8192" "2048 class" "1440p"}; do
 for v in "NVPYR_CASCADE=0"; do
  echo "== $c $v"; env $v NVPYR_TAIL_DEBUG_CLOCKS=1 tools/bench_native --batches 1 --only "$c" 2>&1 | grep "tail clocks" | tail -4
 done
done
