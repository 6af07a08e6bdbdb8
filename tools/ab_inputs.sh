#!/bin/bash
# tools/ab_inputs.sh lib1.so lib2.so ... : 16384^2 chain + 6-level kernel on the three bench inputs, per library build
for lib in "$@"; do
  NVPYR_LIB_PATH=$PWD/vk_compute_mipmaps_b200/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-batch --no-e2e 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', ' | '.join('%s chain %.1f (%.3f) kernel %.1f (%.3f)' % (n, v['us_per_chain'], v['frac_of_hbm_peak'], v['kernel_us'], v['kernel_frac_of_hbm_peak']) for n, v in d['inputs'].items()))"
done
