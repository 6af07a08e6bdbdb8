O=gpurun_out
ncu --set full --import-source on --clock-control none -k regex:fastSrgba8Kernel --launch-skip 3 -c 1 -f -o $O/r2c_fast6_julia \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-inputs --no-batch --no-e2e --input julia > $O/r2c_ncu.log 2>&1
tail -3 $O/r2c_ncu.log
ls -la $O/*.ncu-rep
