#!/bin/bash
# A/B of library builds on the config table: tools/ab_cfg.sh "<only-filter>" lib1.so lib2.so ...
F=$1; shift
for lib in "$@"; do echo "== $lib"; NVPYR_LIB_PATH=$PWD/vk_compute_mipmaps_b200/$lib python tools/bench_configs.py --only "$F" --batches 20 2>&1 | grep -v "^$"; done
