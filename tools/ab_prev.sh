#!/bin/bash
# A/B of two library builds on the native config table, same box: libnvpyr_prev.so (LD_PRELOAD) against libnvpyr.so.
for rep in 1 2; do
  echo "== prev"; LD_PRELOAD=$PWD/vk_compute_mipmaps_b200/libnvpyr_prev.so tools/bench_native --batches 30 2>&1 | cut -c1-100
  echo "== new";  tools/bench_native --batches 30 2>&1 | cut -c1-100
done
