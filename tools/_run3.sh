O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/t_default.log 2>&1; tail -3 $O/t_default.log
L="libnvpyr_w8s16.so libnvpyr_u1c.so libnvpyr_u1.so libnvpyr.so libnvpyr_u2c.so"
bash tools/ab_libs.sh random $L 2>&1 | grep chain_us | tee $O/ab_random.txt
bash tools/ab_libs.sh julia $L 2>&1 | grep chain_us | tee $O/ab_julia.txt
for v in "" _u1c; do
  NVPYR_LIB_PATH=$PWD/vk_compute_mipmaps_b200/libnvpyr$v.so ncu --set full --import-source on --clock-control none -k regex:fastSrgba8Kernel --launch-skip 3 -c 1 -o $O/fast6_julia$v \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-inputs --no-batch --input julia > $O/ncu_fast6$v.log 2>&1
  python tools/ncu_summary.py $O/fast6_julia$v.ncu-rep $O/fast6_julia${v}_summary.txt
done
ls -la $O
