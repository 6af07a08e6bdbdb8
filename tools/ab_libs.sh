#!/bin/bash
# tools/ab_libs.sh <input> lib1.so lib2.so ... : headline kernel timing of several builds on one box
IN=$1; shift
for rep in 1 2; do for lib in "$@"; do
  NVPYR_LIB_PATH=$PWD/vk_compute_mipmaps_b200/$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-inputs --input $IN 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib $IN chain_us %.1f kernel_us %.1f frac %.3f' % (1e3*d['ms_per_step'], d['roofline']['us_per_launch'], d['roofline']['frac']))"
done; done
