#!/bin/bash
# Host round trip: levels downloaded band by band (NVPYR_HOST_BAND_LEVELS) and the tapered last band (NVPYR_HOST_TAPER), one box.
for rep in 1 2; do for v in "NVPYR_HOST_BAND_LEVELS=2 NVPYR_HOST_TAPER=0" "NVPYR_HOST_TAPER=0" "NVPYR_HOST_TAPER=1"; do
  env $v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-inputs --no-batch --input julia 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$v e2e_ms %.3f GB/s %.2f separate_ms %.3f pageable_ms %.3f h2d_only_ms %.3f' % (e['ms_per_step'], e['value'], e['separate_buffers']['ms_per_step'], e['pageable_caller']['ms_per_step'], e['h2d_only']['ms']))"
done; done
