for rep in 1 2; do for v in 0 1 2; do echo "== mode$v"; NVPYR_TAIL_DEFER_WAIT=$v tools/bench_native --batches 30 2>&1 | cut -c1-100; done; done
