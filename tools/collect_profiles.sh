#!/bin/bash
# Round-end evidence, one GPU: bench line, config table, ncu launch list of the bench command, ncu --set full of the
# dominant kernel (Julia and random input) and of the general kernel.  Outputs under gpurun_out/.
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err
python tools/bench_configs.py --batches 30 --out $O/configs.json > $O/configs.txt 2>&1
python tools/bench_batch.py --textures 128 --out $O/batch128.json > $O/batch.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_bench.csv \
    -k regex:'fastSrgba8Kernel|tailKernel|tailBatchKernel|generalSrgba8Kernel|fastKernel|generalKernel|premultiplyKernel' \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
for inp in julia random; do
  ncu --set full --import-source on --clock-control none -k regex:fastSrgba8Kernel --launch-skip 3 -c 1 -o $O/fast6_$inp \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-inputs --no-batch --input $inp > $O/ncu_fast6_$inp.log 2>&1
done
ncu --set full --import-source on --clock-control none -k regex:generalStrip4Kernel -c 1 -o $O/gen4095 \
    python tools/launch_probe.py --only 4095.jpg --reps 1 > $O/ncu_gen.log 2>&1
ls -la $O
