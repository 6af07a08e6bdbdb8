#!/bin/bash
# Round evidence, one GPU: bench line, config table, batch table, ncu launch list of the bench command, ncu --set full of
# the dominant kernel (Julia and random input) and of the general kernel.  Outputs under gpurun_out/ with prefix $1.
set -x
O=gpurun_out
P=${1:-r2_v3}
python bench.py --steps 20 --warmup 3 > $O/${P}_bench.json 2> $O/${P}_bench.err
python tools/bench_configs.py --batches 30 --out $O/${P}_configs.json > $O/${P}_configs.txt 2>&1
python tools/bench_batch.py --textures 128 --out $O/${P}_batch128_4096.json > $O/${P}_batch.txt 2>&1
python tools/warm_launches.py > $O/${P}_warm_launches.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/${P}_launches_bench_steps2.csv \
    -k regex:'fastSrgba8Kernel|tailKernel|tailBatchKernel|generalStrip|fastKernel|generalKernel|premultiply|blitKernel' \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${P}_bench_under_ncu.log 2>&1
for inp in julia random; do
  ncu --set full --import-source on --clock-control none -k regex:fastSrgba8Kernel --launch-skip 3 -c 1 -f -o $O/${P}_fast6_$inp \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-inputs --no-batch --no-e2e --input $inp > $O/${P}_ncu_fast6_$inp.log 2>&1
done
ncu --set full --import-source on --clock-control none -k regex:generalStrip4Kernel --launch-skip 2 -c 1 -f -o $O/${P}_gen4095 \
    python tools/launch_probe.py --only 4095.jpg --reps 4 > $O/${P}_ncu_gen.log 2>&1
ls -la $O | tail -20
