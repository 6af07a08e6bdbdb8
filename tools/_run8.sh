O=gpurun_out
(NVPYR_SLAB_MAX_TILES_PER_WARP_X100=1000000 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > $O/r2h_gputests_slab.txt 2>&1
cat $O/r2h_gputests_slab.txt
for t in 25 90 200 400 800 2000; do
  echo "== NVPYR_SLAB_MAX_TILES_PER_WARP_X100=$t"
  for c in 4096.jpg 8192 2048 1440p 4k 16384; do
  NVPYR_SLAB_MAX_TILES_PER_WARP_X100=$t python tools/bench_configs.py --only "$c" --batches 20 2>&1 | grep -v "^$"
  done
done 2>&1 | tee $O/r2h_slab_sweep.txt
