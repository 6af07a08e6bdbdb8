O=gpurun_out
NVPYR_SLAB_MAX_TILES_PER_WARP_X100=200 ncu --set full --import-source on --clock-control none -k regex:fastSrgba8Kernel --launch-skip 2 -c 1 -f -o $O/r2i_fast6_4096_slab python tools/launch_probe.py --only 4096.jpg --reps 4 > $O/r2i_ncu.log 2>&1
NVPYR_SLAB_MAX_TILES_PER_WARP_X100=25 ncu --set full --import-source on --clock-control none -k regex:fastSrgba8Kernel --launch-skip 2 -c 1 -f -o $O/r2i_fast6_4096_tile python tools/launch_probe.py --only 4096.jpg --reps 4 >> $O/r2i_ncu.log 2>&1
tail -3 $O/r2i_ncu.log
