#!/bin/bash
# A/B of the cascade tail (tools/ab_cascade.sh): the config table with the cascade's knobs varied, on one box.
NVPYR_CASCADE_DEBUG=1 $PRE tools/bench_native --batches 1 2>&1 | grep "nvpyr cascade" | sort | uniq -c
for v in "NVPYR_CASCADE=1" "NVPYR_CASCADE=0" $EXTRA; do echo "== $v"; env $v tools/bench_native --batches 30 2>&1 | grep -v "16384\|8192\|4096.jpg\|2048 class\|4096 rgba" | cut -c1-95; done
